from .allegro_models import (  # noqa: F401
    AllegroEnergyModel,
    AllegroModel,
    ForceStressOutput,
    FullAllegroEnergyModel,
    FullAllegroModel,
    FusedAllegroEnergy,
)
