"""
MRTM river routing on the B200 - drop-in for xanthos/routing/mrtm.py.

Same module-level functions and argument meaning as the reference:
`downstream`, `upstream`, `upstream_genmatrix`, `streamrouting`
(called from Components.calculate_routing, xanthos/components.py:268-289),
plus `route`, which runs the whole spin-up + simulation loop of
components.py:273-294 in ONE persistent kernel launch.

The integer topology is computed by the library's host code (bit-exact with the
reference); the time stepping runs in hand-written CUDA (csrc/mrtm.cu).
"""

import ctypes

import numpy as np

from .. import _cuda as C

_I64P = ctypes.POINTER(ctypes.c_int64)


def _grid_shape(settings):
    return int(settings.ngridrow), int(settings.ngridcol)


def downstream(coord, flowdir, settings):
    """Downstream cell id per cell (1-based, -1 = outlet); reference mrtm.py:85-120."""
    nrow, ncol = _grid_shape(settings)
    co, cop = C.as_c(coord, np.float64)
    fd, fdp = C.as_c(flowdir, np.float64)
    n = co.shape[0]
    dsid = np.zeros(n, dtype=np.int64)
    C.check(C.lib().xan_mrtm_downstream(cop, fdp, n, nrow, ncol, dsid.ctypes.data_as(_I64P)))
    return dsid


def upstream(coord, downstream, settings):
    """[N, 9] neighbour ids, inflowing first, count in column 8; reference mrtm.py:123-191."""
    nrow, ncol = _grid_shape(settings)
    co, cop = C.as_c(coord, np.float64)
    ds, dsp = C.as_c(downstream, np.int64)
    n = co.shape[0]
    upid = np.zeros((n, 9), dtype=np.int64)
    C.check(C.lib().xan_mrtm_upstream(cop, dsp, n, nrow, ncol, upid.ctypes.data_as(_I64P)))
    return upid


class UpstreamMatrix:
    """
    What `upstream_genmatrix` returns: the operator UM = UP - I (reference mrtm.py:194-230) held as
    an execution plan of the CUDA library instead of a scipy sparse matrix.

    `tocsr()` materialises the identical scipy CSR matrix (int64 data, sorted indices) for
    inspection; `streamrouting` / `route` consume the plan directly.
    """

    def __init__(self, upid, block_threads=0, chunk_substeps=0):
        up, upp = C.as_c(upid, np.int64)
        self.shape = (up.shape[0], up.shape[0])
        self.ncell = up.shape[0]
        self._plan = C.check_ptr(C.lib().xan_mrtm_plan_create(upp, self.ncell, block_threads, chunk_substeps))

    def __del__(self):
        try:
            if getattr(self, '_plan', None):
                C.lib().xan_mrtm_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    @property
    def info(self):
        buf = (ctypes.c_int * 8)()
        C.check(C.lib().xan_mrtm_plan_info(self._plan, buf))
        keys = ('is_forest', 'n_components', 'max_component', 'n_warps', 'n_cut_edges', 'n_levels',
                'block_threads', 'max_ghosts')
        return dict(zip(keys, list(buf)))

    def csr_arrays(self):
        nnz = C.lib().xan_mrtm_plan_um_nnz(self._plan)
        indptr = np.zeros(self.ncell + 1, dtype=np.int64)
        indices = np.zeros(nnz, dtype=np.int64)
        data = np.zeros(nnz, dtype=np.int64)
        C.check(C.lib().xan_mrtm_plan_um(self._plan, indptr.ctypes.data_as(_I64P), indices.ctypes.data_as(_I64P),
                                         data.ctypes.data_as(_I64P)))
        return indptr, indices, data

    def tocsr(self):
        import scipy.sparse as sparse
        indptr, indices, data = self.csr_arrays()
        return sparse.csr_matrix((data, indices, indptr), shape=self.shape)

    def packing(self):
        """(lane_cell [n_warps, 32], edge_prod, edge_cons) of the warp kernel, for tests."""
        i = self.info
        nb, c = i['n_warps'], 32
        slot = np.full(max(nb * c, 1), -1, dtype=np.int32)
        ep = np.zeros(max(i['n_cut_edges'], 1), dtype=np.int32)
        ec = np.zeros(max(i['n_cut_edges'], 1), dtype=np.int32)
        ip = ctypes.POINTER(ctypes.c_int)
        C.check(C.lib().xan_mrtm_plan_packing(self._plan, slot.ctypes.data_as(ip), ep.ctypes.data_as(ip),
                                              ec.ctypes.data_as(ip)))
        return slot[:nb * c].reshape(nb, c), ep[:i['n_cut_edges']], ec[:i['n_cut_edges']]


def upstream_genmatrix(upid, block_threads=0, chunk_substeps=0):
    """UM = UP - I as a routing plan; reference mrtm.py:194-230."""
    return UpstreamMatrix(upid, block_threads, chunk_substeps)


def route_device(um, runoff, flow_dist, str_velocity, area, ndays, dt, spinup_months, chs_prev=None,
                 method=C.MRTM_AUTO, want_chs=True, want_avg=True):
    """
    Spin-up + simulation on device.  `runoff` is a Field (or anything `as_field` accepts);
    returns (ChStorage Field, Avg_ChFlow Field, instream_flow cuda tensor [ncell]).
    """
    torch = C.torch_cuda()
    q = C.as_field(runoff)
    n, m = q.ncell, q.nmonths
    L = C.dev_vector(flow_dist)
    V = C.dev_vector(str_velocity)
    A = C.dev_vector(area)
    S0 = None if chs_prev is None else C.dev_vector(chs_prev)
    nd, ndp = C.as_c(np.asarray(ndays).reshape(-1)[:m], np.int32)
    chs = C.Field.empty(n, m, q.ld) if want_chs else None
    avg = C.Field.empty(n, m, q.ld) if want_avg else None
    inst = torch.empty(n, dtype=torch.float64, device='cuda')
    C.check(C.lib().xan_mrtm_route(um._plan, C.ptr(q.t), C.ptr(L), C.ptr(V), C.ptr(A), C.ptr(S0), ndp, m,
                                   int(spinup_months), q.ld, float(dt), int(method),
                                   C.ptr(chs.t if chs else None), C.ptr(avg.t if avg else None), C.ptr(inst),
                                   C.stream_ptr()))
    return chs, avg, inst


def route_device_batch(um, runoffs, flow_dist, str_velocity, area, ndays, dt, spinup_months, method=C.MRTM_AUTO,
                       want_chs=True, want_avg=True):
    """
    Ensemble routing: `runoffs` is a list of Fields (one per member, same shape and leading dimension).
    Returns a list of (ChStorage Field, Avg_ChFlow Field, instream_flow tensor), bit-identical to one
    `route_device` call per member.  On a river forest two members share one launch of the skew kernel: their thread
    blocks are co-resident on every SM and fill each other's idle issue slots (32 instead of 43 ms per member on the
    0.5 degree world; `XANTHOS_MRTM_SKEW_MEMBERS=1` routes one member per launch).
    """
    import ctypes
    torch = C.torch_cuda()
    qs = [C.as_field(r) for r in runoffs]
    n, m, ld = qs[0].ncell, qs[0].nmonths, qs[0].ld
    if any((q.ncell, q.nmonths, q.ld) != (n, m, ld) for q in qs):
        raise C.ValidationException("route_device_batch: members differ in shape")
    L, V, A = C.dev_vector(flow_dist), C.dev_vector(str_velocity), C.dev_vector(area)
    nd, ndp = C.as_c(np.asarray(ndays).reshape(-1)[:m], np.int32)
    k = len(qs)
    outs = [(C.Field.empty(n, m, ld) if want_chs else None, C.Field.empty(n, m, ld) if want_avg else None,
             torch.empty(n, dtype=torch.float64, device='cuda')) for _ in range(k)]
    arr = ctypes.c_void_p * k

    def ptrs(ts):
        return arr(*[None if t is None else t.data_ptr() for t in ts])
    C.check(C.lib().xan_mrtm_route_batch(um._plan, k, ptrs([q.t for q in qs]), C.ptr(L), C.ptr(V), C.ptr(A), None, ndp,
                                         m, int(spinup_months), ld, float(dt), int(method),
                                         ptrs([o[0].t if o[0] else None for o in outs]),
                                         ptrs([o[1].t if o[1] else None for o in outs]),
                                         ptrs([o[2] for o in outs]), C.stream_ptr()))
    return outs


def route(um, runoff, flow_dist, str_velocity, area, ndays, dt, spinup_months, chs_prev=None, method=C.MRTM_AUTO):
    """
    The routing loops of Components.calculate_routing (components.py:273-294) in one call.
    runoff [N, M] (host array, or the array a previous stage returned); returns host arrays
    (ChStorage [N, M], Avg_ChFlow [N, M], instream_flow [N]).
    """
    chs, avg, inst = route_device(um, runoff, flow_dist, str_velocity, area, ndays, dt, spinup_months, chs_prev,
                                  method)
    chs_h = C.remember(chs.to_host(), chs)
    avg_h = C.remember(avg.to_host(), avg)
    return chs_h, avg_h, inst.cpu().numpy()


def streamrouting(L, S0, F0, ChV, q, area, nday, dt, UM):
    """
    One month of routing; same signature and return values (S, Favg, F) as reference mrtm.py:16-82.
    `UM` is the object returned by `upstream_genmatrix`.  F0 is unused, as in the reference.
    """
    torch = C.torch_cuda()
    n = UM.ncell
    qf = torch.from_numpy(np.ascontiguousarray(np.asarray(q, dtype=np.float64).reshape(1, n))).cuda()
    ld = C.padded_ld(n)
    qpad = torch.zeros((1, ld), dtype=torch.float64, device='cuda')
    qpad[:, :n] = qf
    chs, avg, inst = route_device(UM, C.Field(qpad, n), L, ChV, area, [int(nday)], dt, 0, chs_prev=S0)
    return chs.t[0, :n].cpu().numpy(), avg.t[0, :n].cpu().numpy(), inst.cpu().numpy()
