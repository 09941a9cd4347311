from .common import IncompressibleEuler
from .hdg_implicit import IncompressibleEulerHDGImplicit
from .hdg_imex import (
    IncompressibleEulerHDGIMEX,
    IncompressibleEulerHDGIMEXImplicit,
    IncompressibleEulerHDGIMEXARS2_232,
    IncompressibleEulerHDGIMEXARS3_443,
    IncompressibleEulerHDGIMEXSSP2_332,
    IncompressibleEulerHDGIMEXSSP3_433,
)

__all__ = [
    "IncompressibleEuler",
    "IncompressibleEulerHDGImplicit",
    "IncompressibleEulerHDGIMEX",
    "IncompressibleEulerHDGIMEXImplicit",
    "IncompressibleEulerHDGIMEXARS2_232",
    "IncompressibleEulerHDGIMEXARS3_443",
    "IncompressibleEulerHDGIMEXSSP2_332",
    "IncompressibleEulerHDGIMEXSSP3_433",
]
