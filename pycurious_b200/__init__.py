"""pycurious_b200 -- B200-native per-window Curie-depth pipeline behind the pycurious API.

Drop-in for the hot path of brmather/pycurious (window -> radial spectrum -> Bouligand MAP fit
-> Metropolis-Hastings): `CurieGrid`, `CurieOptimise`, `bouligand2009`, `maus1995`,
`tanaka1999`, `ComputeTanaka` keep the reference's names and signatures
(pycurious/__init__.py:145-149).  Compute runs in hand-written sm_100a CUDA kernels
(pycurious_b200/csrc) through the C-ABI in include/curie_b200.h.  Mapping, download and
documentation helpers of the reference are out of scope.
"""
from .grid import CurieGrid, bouligand2009, tanaka1999, maus1995, ComputeTanaka
from .optimise import CurieOptimise
from . import radon
from . import distributed
from . import synth

__version__ = "0.1.0"
__all__ = ["CurieGrid", "CurieOptimise", "bouligand2009", "tanaka1999", "maus1995", "ComputeTanaka",
           "radon", "distributed", "synth"]
