"""Thin Python holder of one gsm_handle (one chain / one batch) — plumbing only: it moves pointers
across the C ABI of include/gsm_b200.h and never computes a branch term itself."""
import ctypes as C

import numpy as np

from . import _abi as abi
from . import _lib


class GsmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gsm_b200 error %d: %s" % (code, msg))
        self.code = code


def _np(a, dtype, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (shape, a.shape))
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dp(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else C.c_void_p(t.data_ptr())


class GsmEngine:
    def __init__(self, max_trees=0, max_nodes=0, device=0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.gsm_create(C.byref(self._h), int(device), int(max_trees), int(max_nodes))
        if rc != abi.OK:
            raise GsmError(rc, (self._lib.gsm_last_error(None) or b"").decode())
        self.device = int(device)
        self.n_trees = 0
        self.n_nodes = 0
        self._keep = None

    def close(self):
        if self._h:
            self._lib.gsm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != abi.OK:
            raise GsmError(rc, (self._lib.gsm_last_error(self._h) or b"").decode())

    # ---- wiring ----
    def set_mode(self, stub_mode, flags):
        self._ck(self._lib.gsm_set_mode(self._h, int(stub_mode), int(flags)))

    def set_option(self, name, value):
        self._ck(self._lib.gsm_set_option(self._h, name.encode(), int(value)))

    # ---- host-buffer path ----
    def upload_trees(self, height, parent, node_count=None):
        height = _np(height, np.float64)
        T, N = height.shape
        parent = _np(parent, np.int32, (T, N))
        node_count = _np(node_count, np.int32, (T,))
        self._ck(self._lib.gsm_upload_trees(self._h, T, N, _p(height), _p(parent), _p(node_count)))
        self.n_trees, self.n_nodes = T, N

    def upload_branch_params(self, rate=None, spike=None, nstubs=None):
        shp = (self.n_trees, self.n_nodes)
        rate, spike, nstubs = _np(rate, np.float64, shp), _np(spike, np.float64, shp), _np(nstubs, np.int32, shp)
        self._ck(self._lib.gsm_upload_branch_params(self._h, _p(rate), _p(spike), _p(nstubs)))

    def upload_tree_params(self, params, use_spike=None):
        params = _np(params, np.float64, (self.n_trees, abi.NPARAM))
        use_spike = _np(use_spike, np.uint8, (self.n_trees,))
        self._ck(self._lib.gsm_upload_tree_params(self._h, _p(params), _p(use_spike)))

    def upload_stubs(self, stub_ptr, stub_branch, stub_time):
        """EXACT mode stub lists, CSR by tree (gsm_upload_stubs)."""
        stub_ptr = _np(stub_ptr, np.int64, (self.n_trees + 1,))
        S = int(stub_ptr[-1])
        stub_branch = _np(stub_branch, np.int32, (S,))
        stub_time = _np(stub_time, np.float64, (S,))
        self._ck(self._lib.gsm_upload_stubs(self._h, _p(stub_ptr), _p(stub_branch), _p(stub_time)))

    def bind_stubs(self, stub_ptr, stub_branch, stub_time):
        """device tensors (int64 / int32 / float64) of the stub lists; kept alive by the engine"""
        self._keep_stubs = (stub_ptr, stub_branch, stub_time)
        self._ck(self._lib.gsm_bind_device_stubs(self._h, _dp(stub_ptr), _dp(stub_branch), _dp(stub_time)))

    def eval(self, want_rates=False, want_wspikes=False):
        T, N = self.n_trees, self.n_nodes
        out = np.empty((T, abi.NOUT), np.float64)
        rates = np.empty((T, N), np.float64) if want_rates else None
        wsp = np.empty((T, N), np.float64) if want_wspikes else None
        self._ck(self._lib.gsm_eval(self._h, _p(out), _p(rates), _p(wsp)))
        return {"out": out, "rates": rates, "wspikes": wsp}

    def eval_host(self, height, parent, spike, params, rate=None, nstubs=None, use_spike=None, node_count=None,
                  want_rates=False, want_wspikes=False, out=None):
        height = _np(height, np.float64)
        T, N = height.shape
        parent = _np(parent, np.int32, (T, N))
        spike = _np(spike, np.float64, (T, N))
        rate = _np(rate, np.float64, (T, N))
        nstubs = _np(nstubs, np.int32, (T, N))
        params = _np(params, np.float64, (T, abi.NPARAM))
        use_spike = _np(use_spike, np.uint8, (T,))
        node_count = _np(node_count, np.int32, (T,))
        if out is None:
            out = np.empty((T, abi.NOUT), np.float64)
        rates = np.empty((T, N), np.float64) if want_rates else None
        wsp = np.empty((T, N), np.float64) if want_wspikes else None
        self._ck(self._lib.gsm_eval_host(self._h, T, N, _p(height), _p(parent), _p(node_count), _p(rate), _p(spike),
                                         _p(nstubs), _p(params), _p(use_spike), _p(out), _p(rates), _p(wsp)))
        return {"out": out, "rates": rates, "wspikes": wsp}

    def eval_host_ptrs(self, T, N, height, parent, node_count, rate, spike, nstubs, params, use_spike, out_tree,
                       out_rates=None, out_wspikes=None):
        """Raw-pointer form (ints / c_void_p), e.g. pinned torch tensors' data_ptr()."""
        vp = lambda x: None if x is None else C.c_void_p(int(x))
        self._ck(self._lib.gsm_eval_host(self._h, int(T), int(N), vp(height), vp(parent), vp(node_count), vp(rate),
                                         vp(spike), vp(nstubs), vp(params), vp(use_spike), vp(out_tree), vp(out_rates),
                                         vp(out_wspikes)))

    def eval_dirty(self, tree_of_prop, dirty_ptr, dirty_nodes, new_height=None, new_rate=None, new_spike=None,
                   new_nstubs=None, height_scale=None, new_params=None, new_use_spike=None, out=None):
        """Proposal batch on the resident batch (gsm_eval_dirty): returns out [n_props][NOUT]."""
        tree_of_prop = _np(tree_of_prop, np.int32)
        P = tree_of_prop.shape[0]
        dirty_ptr = _np(dirty_ptr, np.int64, (P + 1,))
        D = int(dirty_ptr[-1]) if P else 0
        dirty_nodes = _np(dirty_nodes, np.int32, (D,))
        new_height = _np(new_height, np.float64, (D,))
        new_rate = _np(new_rate, np.float64, (D,))
        new_spike = _np(new_spike, np.float64, (D,))
        new_nstubs = _np(new_nstubs, np.int32, (D,))
        height_scale = _np(height_scale, np.float64, (P,))
        new_params = _np(new_params, np.float64, (P, abi.NPARAM))
        new_use_spike = _np(new_use_spike, np.uint8, (P,))
        if out is None:
            out = np.empty((P, abi.NOUT), np.float64)
        self._ck(self._lib.gsm_eval_dirty(self._h, P, _p(tree_of_prop), _p(dirty_ptr), _p(dirty_nodes), _p(new_height),
                                          _p(new_rate), _p(new_spike), _p(new_nstubs), _p(height_scale), _p(new_params),
                                          _p(new_use_spike), _p(out)))
        return out

    @property
    def last_dirty_full_count(self):
        return int(self._lib.gsm_last_dirty_full_count(self._h))

    def stub_cdf(self, tree, node, cap=4096):
        buf = np.empty(cap, np.float64)
        n = C.c_int32(0)
        self._ck(self._lib.gsm_stub_cdf(self._h, int(tree), int(node), _p(buf), cap, C.byref(n)))
        return buf[: min(n.value, cap)].copy(), n.value

    def stub_cdf_trees(self, trees=None, n_sel=None):
        """getCumulativeProbs of every branch of the selected trees: returns (cdf_ptr [n_sel * N + 1], cdf)"""
        trees = _np(trees, np.int64)
        n_sel = len(trees) if trees is not None else (self.n_trees if n_sel is None else int(n_sel))
        ptr = np.zeros(n_sel * self.n_nodes + 1, np.int64)
        need = C.c_int64(0)
        cdf = np.empty(max(1, 16 * n_sel * self.n_nodes), np.float64)
        self._ck(self._lib.gsm_stub_cdf_trees(self._h, n_sel, _p(trees), _p(ptr), _p(cdf), cdf.size, C.byref(need)))
        if need.value > cdf.size:
            cdf = np.empty(need.value, np.float64)
            self._ck(self._lib.gsm_stub_cdf_trees(self._h, n_sel, _p(trees), _p(ptr), _p(cdf), cdf.size, C.byref(need)))
        return ptr, cdf[: need.value].copy()

    def stub_choose(self, u, trees=None):
        """Stubs.sampleNStubsOnBranch for every branch of the selected trees given uniforms u [n_sel][N]"""
        u = _np(u, np.float64)
        n_sel = u.shape[0]
        trees = _np(trees, np.int64, (n_sel,))
        out = np.empty((n_sel, self.n_nodes), np.int32)
        self._ck(self._lib.gsm_stub_choose(self._h, n_sel, _p(trees), _p(u), _p(out)))
        return out

    # ---- device-resident path (torch tensors own the memory) ----
    def bind_batch(self, batch):
        """batch: synthetic.Batch (or anything with the same tensor attributes) already on this device."""
        self._keep = batch
        self._ck(self._lib.gsm_bind_device(self._h, batch.n_trees, batch.n_nodes, batch.stride, _dp(batch.height),
                                           _dp(batch.parent), _dp(getattr(batch, "node_count", None)), _dp(batch.rate),
                                           _dp(batch.spike), _dp(batch.nstubs), _dp(batch.params), _dp(batch.use_spike)))
        self.n_trees, self.n_nodes = batch.n_trees, batch.n_nodes

    def eval_device(self, out_tree, out_rates=None, out_wspikes=None, stream=None):
        """Asynchronous launch on `stream` (an int cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream)."""
        self._ck(self._lib.gsm_eval_device(self._h, _dp(out_tree), _dp(out_rates), _dp(out_wspikes),
                                           C.c_void_p(stream) if stream else None))

    def measure_fp64_peak(self):
        """thread-level DFMA per second of this device (micro-benchmark; the denominator of bench.py's roofline.fp64)"""
        v = C.c_double(0.0)
        self._ck(self._lib.gsm_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def sync(self):
        self._ck(self._lib.gsm_sync(self._h))

    @property
    def launch_count(self):
        return int(self._lib.gsm_launch_count(self._h))
