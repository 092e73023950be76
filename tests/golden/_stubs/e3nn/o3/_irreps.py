"""e3nn.o3._irreps look-alike: only the bookkeeping API the reference touches
(allegro/nn/_allegro.py:58-160, allegro/nn/_strided/_contract.py:48-118)."""
from oracle.o3_ref import Irrep  # (l, p) with .l .p .dim, ir * ir -> [Irrep], "1o" parsing


class _MulIr(tuple):
    def __new__(cls, mul, ir):
        return super().__new__(cls, (int(mul), Irrep(ir)))

    @property
    def mul(self):
        return self[0]

    @property
    def ir(self):
        return self[1]

    @property
    def dim(self):
        return self[0] * self[1].dim


class Irreps(tuple):
    def __new__(cls, spec=None):
        if isinstance(spec, Irreps):
            return spec
        out = []
        if spec is None:
            spec = []
        if isinstance(spec, Irrep):
            spec = [(1, spec)]
        if isinstance(spec, str):
            for term in filter(None, (t.strip() for t in spec.split("+"))):
                if "x" in term:
                    mul, ir = term.split("x")
                    out.append(_MulIr(int(mul), Irrep(ir)))
                else:
                    out.append(_MulIr(1, Irrep(term)))
        else:
            for item in spec:
                if isinstance(item, (Irrep, str)):
                    out.append(_MulIr(1, Irrep(item)))
                else:
                    mul, ir = item
                    out.append(_MulIr(mul, Irrep(ir)))
        return super().__new__(cls, out)

    @staticmethod
    def spherical_harmonics(lmax, p=-1):
        return Irreps([(1, (l, p**l)) for l in range(lmax + 1)])

    @property
    def dim(self):
        return sum(mi.dim for mi in self)

    @property
    def num_irreps(self):
        return sum(mi.mul for mi in self)

    @property
    def lmax(self):
        return max(mi.ir.l for mi in self)

    @property
    def ls(self):
        return [mi.ir.l for mi in self for _ in range(mi.mul)]

    def slices(self):
        out, i = [], 0
        for mi in self:
            out.append(slice(i, i + mi.dim))
            i += mi.dim
        return out

    def __contains__(self, ir):
        ir = Irrep(ir)
        return any(mi.ir == ir for mi in self)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return Irreps(tuple.__getitem__(self, i))
        return tuple.__getitem__(self, i)

    def __add__(self, other):
        return Irreps(tuple(self) + tuple(Irreps(other)))

    def __repr__(self):
        return "+".join(f"{mi.mul}x{mi.ir}" for mi in self)
