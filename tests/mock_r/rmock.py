"""Python side of the R stand-in (tests/mock_r): builds r-package/src/hb_rglue.c against the mock R API and offers
`.Call`-like access to its registered entry points.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_build", "libhb_rglue_mock.so")
NILSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP, EXTPTRSXP = 0, 10, 13, 14, 16, 19, 22


class RError(RuntimeError):
    """error() inside a .Call -- what R would raise as a condition (stop())."""


def build() -> str:
    srcs = [os.path.join(ROOT, "r-package", "src", "hb_rglue.c"), os.path.join(HERE, "mock_r.c")]
    deps = srcs + [os.path.join(HERE, "include", "Rinternals.h"), os.path.join(ROOT, "include", "hydrobayes_b200.h"),
                   os.path.join(ROOT, "hydrobayes_b200", "libhydrobayes_b200.so")]
    if not os.path.exists(SO) or any(os.path.getmtime(SO) < os.path.getmtime(p) for p in deps):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        libdir = os.path.join(ROOT, "hydrobayes_b200")
        subprocess.run(["gcc", "-O1", "-Wall", "-Werror=implicit-function-declaration", "-shared", "-fPIC",
                        "-I" + os.path.join(HERE, "include"), "-I" + os.path.join(ROOT, "include"), *srcs, "-o", SO,
                        "-L" + libdir, "-lhydrobayes_b200", "-Wl,-rpath," + libdir, "-lm"], check=True)
    return SO


class MockR:
    def __init__(self):
        L = self.L = C.CDLL(build())
        vp = C.c_void_p
        L.mockR_load.restype = C.c_int
        L.mockR_entry_name.restype = C.c_char_p; L.mockR_entry_name.argtypes = [C.c_int]
        L.mockR_entry_nargs.argtypes = [C.c_int]
        L.mockR_dotCall.restype = vp; L.mockR_dotCall.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp), C.c_char_p, C.c_int]
        L.mockR_real.restype = vp; L.mockR_real.argtypes = [C.POINTER(C.c_double), C.c_ssize_t]
        L.mockR_real_na.restype = vp; L.mockR_real_na.argtypes = [C.c_ssize_t]
        L.mockR_int.restype = vp; L.mockR_int.argtypes = [C.POINTER(C.c_int), C.c_ssize_t]
        L.mockR_set_dim.restype = vp; L.mockR_set_dim.argtypes = [vp, C.POINTER(C.c_int), C.c_int]
        L.mockR_list.restype = vp; L.mockR_list.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(vp)]
        L.mockR_nil.restype = vp
        L.mkString.restype = vp; L.mkString.argtypes = [C.c_char_p]
        L.ScalarLogical.restype = vp; L.ScalarLogical.argtypes = [C.c_int]
        L.mockR_type.argtypes = [vp]
        L.mockR_data.restype = vp; L.mockR_data.argtypes = [vp]
        L.XLENGTH.restype = C.c_ssize_t; L.XLENGTH.argtypes = [vp]
        L.VECTOR_ELT.restype = vp; L.VECTOR_ELT.argtypes = [vp, C.c_ssize_t]
        L.mockR_free_all.restype = C.c_int
        self.n_entries = L.mockR_load()                     # R_init_HydroBayes(dll)
        self.nil = L.mockR_nil()

    def entries(self):
        return {self.L.mockR_entry_name(i).decode(): self.L.mockR_entry_nargs(i) for i in range(self.n_entries)}

    # ---- R values ----
    def real(self, v, dim=None):
        """numeric vector / matrix / array (column-major, like R)."""
        a = np.asarray(v, dtype=np.float64)
        if dim is None and a.ndim > 1:
            dim = a.shape
        flat = np.ascontiguousarray(a.reshape(-1, order="F"))
        s = self.L.mockR_real(flat.ctypes.data_as(C.POINTER(C.c_double)), flat.size)
        return self._dim(s, dim)

    def real_na(self, *dim):
        s = self.L.mockR_real_na(int(np.prod(dim)))
        return self._dim(s, dim if len(dim) > 1 else None)

    def _dim(self, s, dim):
        if dim is not None:
            d = (C.c_int * len(dim))(*[int(x) for x in dim])
            self.L.mockR_set_dim(s, d, len(dim))
        return s

    def integer(self, v):
        a = np.ascontiguousarray(np.atleast_1d(np.asarray(v, dtype=np.int32)))
        return self.L.mockR_int(a.ctypes.data_as(C.POINTER(C.c_int)), a.size)

    def string(self, s):
        return self.nil if s is None else self.L.mkString(s.encode())

    def value(self, v):
        if v is None:
            return self.nil
        if isinstance(v, bool):
            return self.L.ScalarLogical(int(v))
        if isinstance(v, str):
            return self.string(v)
        if isinstance(v, (int, np.integer)) and abs(int(v)) < 2 ** 31:
            return self.integer(v)
        return self.real(v)

    def named_list(self, **kw):
        n = len(kw)
        names = (C.c_char_p * n)(*[k.encode() for k in kw])
        vals = (C.c_void_p * n)(*[self.value(v) for v in kw.values()])
        return self.L.mockR_list(n, names, vals)

    # ---- .Call ----
    def call(self, name, *args):
        a = (C.c_void_p * max(len(args), 1))(*args)
        err = C.create_string_buffer(1024)
        res = self.L.mockR_dotCall(name.encode(), len(args), a, err, 1024)
        if res is None:
            raise RError(err.value.decode())
        return res

    # ---- reading results back ----
    def to_numpy(self, s, shape=None):
        t = self.L.mockR_type(s)
        n = self.L.XLENGTH(s)
        ct = {REALSXP: C.c_double, INTSXP: C.c_int, LGLSXP: C.c_int}[t]
        if n == 0:
            return np.empty(0, dtype=ct)
        a = np.ctypeslib.as_array(C.cast(self.L.mockR_data(s), C.POINTER(ct)), (n,)).copy()
        return a if shape is None else a.reshape(shape, order="F")

    def list_elt(self, s, i):
        return self.L.VECTOR_ELT(s, i)

    def free_all(self):
        return self.L.mockR_free_all()
