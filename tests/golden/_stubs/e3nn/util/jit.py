def compile_mode(mode):
    """e3nn.util.jit.compile_mode: only tags the class for TorchScript; no effect in eager mode."""

    def deco(cls):
        cls._e3nn_compile_mode = mode
        return cls

    return deco
