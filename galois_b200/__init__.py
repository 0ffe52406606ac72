"""
galois_b200 — a B200-native (sm_100a) finite-field arithmetic backend with the reference's
(mhostetter/galois) array interface for its data-parallel hot path:

    GF = galois_b200.GF(2**31 - 1)        # np.ndarray subclass, like galois.GF
    x * y, x / y, x ** -1, np.matmul, np.convolve
    galois_b200.ntt(x), galois_b200.intt(X)
    rs = galois_b200.ReedSolomon(255, 223); rs.encode(m); rs.decode(c, errors=True)
    galois_b200.Poly([1, 0, 5], field=GF).roots(); np.linalg.inv(A); A.row_reduce()

Every arithmetic call goes through the C ABI in include/galois_b200.h into hand-written CUDA
kernels (galois_b200/csrc).  There is no CPU fallback.
"""

from ._field import GF, DeviceArray, Field, FieldArray
from ._binding import disable, enable, is_enabled
from ._lib import BackendUnavailable
from ._linalg import poly_evaluate
from ._poly import Poly, poly_roots
from ._ntt import intt, ntt, ntt_device
from ._rs import BCH, ReedSolomon

__all__ = ["GF", "Field", "FieldArray", "DeviceArray", "ntt", "intt", "ntt_device", "ReedSolomon", "BCH", "poly_evaluate", "poly_roots", "Poly",
           "BackendUnavailable", "enable", "disable", "is_enabled"]
__version__ = "0.1.0"
