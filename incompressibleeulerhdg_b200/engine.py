"""ctypes binding of ``libhdg_b200.so`` (the C-ABI declared in ``include/hdg_b200.h``).

Host code stays Python (BASELINE.json north_star); PyTorch is used only to own device buffers and
streams.  There is no CPU fallback: constructing an :class:`HDGEngine` without the CUDA library or
without a GPU raises.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

__all__ = ["HDGEngine", "HDGError", "load_library", "LIB_PATH", "TIMER_LABELS", "nccl_unique_id",
           "broadcast_unique_id"]

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhdg_b200.so")

TIMER_LABELS = ("setup_poisson", "forward_elimination", "trace_solve", "back_substitution", "bdm_projection",
                "tentative_velocity_solve", "h2d", "d2h", "spmv_sampled", "fimpl_sampled", "condense", "assemble")

HDG_OK, HDG_EINVAL, HDG_ECUDA, HDG_ENOGPU, HDG_ESTATE, HDG_ENCCL, HDG_ENOCONV, HDG_ECOMM = range(8)

_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p

class hdg_csr(C.Structure):
    _fields_ = [("nrows", C.c_int32), ("ncols", C.c_int32), ("rowptr", _ip), ("col", _ip), ("val", _dp)]


#: every symbol of include/hdg_b200.h with (restype, argtypes); tests check the library exports all
SIGNATURES = {
    "hdg_version": (C.c_char_p, []),
    "hdg_supported_degrees": (C.c_int, []),
    "hdg_device_count": (C.c_int, []),
    "hdg_create": (C.c_int, [C.c_int, C.c_double, C.c_int, C.c_int, _dp, _ip, _ip, _ip, _ip, C.c_int, C.POINTER(_vp)]),
    "hdg_destroy": (C.c_int, [_vp]),
    "hdg_last_error": (C.c_char_p, [_vp]),
    "hdg_set_stream": (C.c_int, [_vp, _vp]),
    "hdg_synchronize": (C.c_int, [_vp]),
    "hdg_setup_poisson": (C.c_int, [_vp, C.c_int]),
    "hdg_get_local_schur": (C.c_int, [_vp, _dp]),
    "hdg_get_trace_matrix": (C.c_int, [_vp, _dp, _ip]),
    "hdg_poisson_apply_host": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int,
                                         C.POINTER(C.c_int)]),
    "hdg_poisson_apply_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_int, C.c_int,
                                        C.POINTER(C.c_int)]),
    "hdg_poisson_apply_update_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_double, _vp, C.c_double, _vp, C.c_double,
                                               C.c_double, _vp, _vp, C.c_double, C.c_int, C.POINTER(C.c_int)]),
    "hdg_set_initial_guess": (C.c_int, [_vp, C.c_int]),
    "hdg_mg_setup": (C.c_int, [_vp, C.c_int, C.POINTER(hdg_csr), C.POINTER(hdg_csr), C.POINTER(hdg_csr),
                               C.POINTER(hdg_csr), C.POINTER(hdg_csr), _dp, _dp, C.c_int, C.c_int, C.c_double]),
    "hdg_mg_enable": (C.c_int, [_vp, C.c_int]),
    "hdg_mg_info": (C.c_int, [_vp, C.POINTER(C.c_int), _dp]),
    "hdg_trace_spmv_dev": (C.c_int, [_vp, _vp, _vp]),
    "hdg_forward_eliminate_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "hdg_back_substitute_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "hdg_set_penalty": (C.c_int, [_vp, C.c_double]),
    "hdg_set_tentative_solver": (C.c_int, [_vp, C.c_int, C.c_int]),
    "hdg_tentative_stats": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "hdg_mixed_stats": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "hdg_set_tentative_comm": (C.c_int, [_vp, C.c_int]),
    "hdg_project_bdm_dev": (C.c_int, [_vp, _vp, _vp]),
    "hdg_fimpl_apply_dev": (C.c_int, [_vp, _vp, _vp, C.c_double, C.c_double, C.c_int, _vp]),
    "hdg_tentative_solve_dev": (C.c_int, [_vp, _vp, C.c_double, C.c_int, _vp, _vp, C.c_double, C.c_int, C.c_int,
                                          C.POINTER(C.c_int)]),
    "hdg_weak_divergence_dev": (C.c_int, [_vp, _vp, C.c_double, C.c_int, _vp]),
    "hdg_pressure_gradient_dev": (C.c_int, [_vp, _vp, _vp, C.c_double, C.c_double, _vp]),
    "hdg_reconstruct_trace_dev": (C.c_int, [_vp, _vp, _vp, _vp]),
    "hdg_shift_pressure_dev": (C.c_int, [_vp, _vp, _vp]),
    "hdg_reconstruction_rhs_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "hdg_gamma_apply_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "hdg_dot_dev": (C.c_int, [_vp, C.c_int, _vp, _vp, _dp]),
    "hdg_l2_inner_dev": (C.c_int, [_vp, C.c_int, _vp, _vp, _dp]),
    "hdg_lincomb_dev": (C.c_int, [_vp, C.c_int64, _vp, C.c_int, _dp, C.POINTER(_vp)]),
    "hdg_mass_dev": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
    "hdg_tracer_setup": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _ip, _ip, _dp, _dp, C.c_int, _dp, C.c_int, _dp]),
    "hdg_project_cg_dev": (C.c_int, [_vp, _vp, _vp, C.c_double, C.c_int, C.POINTER(C.c_int)]),
    "hdg_tracer_advection_dev": (C.c_int, [_vp, _vp, _vp, C.c_double, _vp, C.c_double, _vp]),
    "hdg_comm_unique_id": (C.c_int, [_vp]),
    "hdg_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "hdg_set_partition": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, C.c_double]),
    "hdg_set_halo_plan": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, _ip, _ip]),
    "hdg_p2p_alloc": (C.c_int, [_vp, C.c_int64, _vp]),
    "hdg_p2p_attach": (C.c_int, [_vp, _vp]),
    "hdg_p2p_enable": (C.c_int, [_vp, C.c_int]),
    "hdg_p2p_status": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "hdg_halo_exchange_dev": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "hdg_allreduce_sum_dev": (C.c_int, [_vp, _vp, C.c_int]),
    "hdg_comm_probe": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp]),
    "hdg_mg_set_distribution": (C.c_int, [_vp, C.c_int, _ip, _ip]),
    "hdg_comm_stats": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int64)]),
    "hdg_field_size": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int64)]),
    "hdg_upload": (C.c_int, [_vp, C.c_int, _dp, _vp]),
    "hdg_download": (C.c_int, [_vp, C.c_int, _vp, _dp]),
    "hdg_get_timers": (C.c_int, [_vp, _dp, C.POINTER(C.c_int64), C.c_int]),
    "hdg_reset_timers": (C.c_int, [_vp]),
    "hdg_measure_fp64_peak": (C.c_int, [_vp, _dp]),
    "hdg_tent_sweep_probe": (C.c_int, [_vp, C.c_double, C.c_int, _dp]),
    "hdg_debug_scalars": (C.c_int, [_vp, _dp]),
    "hdg_set_graphs": (C.c_int, [_vp, C.c_int]),
    "hdg_set_tuning": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "hdg_graph_replays": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "hdg_guess_restarts": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "hdg_launch_count": (C.c_int64, [_vp]),
    "hdg_kernel_counts": (C.c_int64, [_vp, C.c_char_p, C.c_int64]),
    "hdg_kernel_times": (C.c_int64, [_vp, C.c_char_p, C.c_int64]),
    "hdg_upload_begin": (C.c_int, [_vp, C.c_int, _dp, C.c_int]),
    "hdg_upload_end": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "hdg_download_begin": (C.c_int, [_vp, C.c_int, _vp, _dp, C.c_int]),
    "hdg_copy_wait": (C.c_int, [_vp]),
}


class HDGError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hdg_b200 error {code}: {msg}")
        self.code = code


def load_library(path: str | None = None):
    """dlopen the engine; raises if it has not been built (run ``python __graft_entry__.py`` or
    ``python -m incompressibleeulerhdg_b200.build``)"""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise HDGError(HDG_ENOGPU, f"{path} not found: build the CUDA engine first "
                                   "(python -m incompressibleeulerhdg_b200.build); there is no CPU fallback")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def parse_tuning(spec: str):
    """``"tent_cellblock=1, tent_sweeps=4"`` -> ``[("tent_cellblock", 1), ("tent_sweeps", 4)]``; a bare name means 1"""
    out = []
    for item in spec.split(","):
        name, _, value = item.strip().partition("=")
        if not name:
            continue
        try:
            out.append((name.strip(), int(value.strip() or 1)))
        except ValueError:
            raise ValueError(f"HDG_TUNING: {item.strip()!r} is not name=integer") from None
    return out


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _dev(t):
    """raw device pointer of a torch tensor (or None)"""
    if t is None:
        return None
    assert t.is_cuda and t.dtype.is_floating_point and t.element_size() == 8 and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def nccl_unique_id() -> bytes:
    """a fresh 128-byte ncclUniqueId (call on one rank, broadcast to the others)"""
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.hdg_comm_unique_id(buf)
    if rc != HDG_OK:
        raise HDGError(rc, lib.hdg_last_error(None).decode())
    return buf.raw


def broadcast_unique_id() -> bytes:
    """rank 0 creates the ncclUniqueId, torch.distributed carries it to the other ranks"""
    import torch.distributed as dist

    box = [nccl_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


class HDGEngine:
    """One engine per GPU.  Mirrors the role of the SCPC python context of `hdg_imex.py:128-133`.

    `mesh` is either a complete :class:`mesh.Mesh` (single GPU) or a :class:`partition.LocalMesh`
    (one rank of a partitioned mesh; `comm_id` is then the broadcast ncclUniqueId)."""

    def __init__(self, mesh, k: int, tau: float = 1.0, device: int = 0, torch_stream: bool = True,
                 comm_id: bytes | None = None):
        self.lib = load_library()
        if not (self.lib.hdg_supported_degrees() >> int(k)) & 1:
            raise HDGError(HDG_EINVAL, f"degree k={k} is not compiled into {LIB_PATH}")
        self.part = None
        if hasattr(mesh, "cells") and hasattr(mesh, "mesh"):  # partition.LocalMesh
            self.part = mesh
            mesh = mesh.mesh
        self.mesh = mesh
        self.k = int(k)
        self.tau = float(tau)
        self.device = int(device)
        self.nc, self.nf = mesh.nc, mesh.nf
        self.nQ1 = (k + 2) * (k + 3) // 2
        self.np_ = (k + 1) * (k + 2) // 2
        self.nl1 = k + 1
        self._h = _vp()
        xy = np.ascontiguousarray(mesh.cell_xy, dtype=np.float64)
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in
                (mesh.cell_facet, mesh.cell_flip, mesh.facet_cell, mesh.facet_local)]
        rc = self.lib.hdg_create(self.k, self.tau, self.nc, self.nf, _ptr(xy), *[a.ctypes.data_as(_ip) for a in arrs],
                                 self.device, C.byref(self._h))
        if rc != HDG_OK:
            raise HDGError(rc, self.lib.hdg_last_error(None).decode())
        self.last_iterations = 0
        self.p2p = False
        if torch_stream:
            # order engine work with torch's current stream so that torch-owned buffers are safe to share
            self.use_torch_stream()
        if os.environ.get("HDG_GRAPHS", "1") == "0":
            self.set_graphs(False)
        # result-neutral knobs for A/B runs of the unchanged tests / bench, e.g. HDG_TUNING="tent_cellblock=1,tent_sweeps=4"
        self.tuning = {}
        for name, value in parse_tuning(os.environ.get("HDG_TUNING", "")):
            if name == "tent_sweeps":  # Chebyshev sweeps on the facet Schur complement of the tentative solve
                self.set_tentative_solver(1, value)
            elif name == "tent_local_sweeps":  # multi-GPU: no halo exchange between those sweeps
                self.set_tentative_comm(bool(value))
            else:
                self.set_tuning(name, value)
            self.tuning[name] = value
        if self.part is not None and self.part.nranks > 1:
            if comm_id is None:
                comm_id = broadcast_unique_id()
            self.comm_init(self.part.rank, self.part.nranks, comm_id)
            pt = self.part
            self._check(self.lib.hdg_set_partition(self._h, pt.nc_owned, pt.nf_owned, pt.global_nf, pt.global_volume))
            self.set_halo_plan(0, pt.cells)
            self.set_halo_plan(1, pt.facets)
            if os.environ.get("HDG_P2P", "1") != "0":
                self.p2p_setup_or_nccl()

    # -- multi-GPU ----------------------------------------------------------------------------------
    @property
    def nranks(self):
        return 1 if self.part is None else self.part.nranks

    @property
    def rank(self):
        return 0 if self.part is None else self.part.rank

    def comm_init(self, rank: int, nranks: int, comm_id: bytes):
        assert len(comm_id) == 128
        buf = C.create_string_buffer(comm_id, 128)
        self._check(self.lib.hdg_comm_init(self._h, int(rank), int(nranks), buf))

    def set_halo_plan(self, kind: int, plan):
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        peers, sptr, sidx, roff, rcnt = (i32(a) for a in (plan.peers, plan.send_ptr, plan.send_idx, plan.recv_off,
                                                           plan.recv_cnt))
        if sidx.size == 0:
            sidx = np.zeros(1, dtype=np.int32)
        p = lambda a: a.ctypes.data_as(_ip)
        self._check(self.lib.hdg_set_halo_plan(self._h, int(kind), int(plan.n_owned), int(plan.n_local), int(peers.size),
                                               p(peers), p(sptr), p(sidx), p(roff), p(rcnt)))

    def p2p_setup_or_nccl(self):
        """peer-memory transport when every rank can use it (one NVSwitch box, <= 8 ranks, peer access and CUDA IPC
        available), decided collectively; otherwise the NCCL send/recv path stays in place"""
        import warnings

        import torch
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            return  # a caller-supplied communicator without a process group: NCCL transport only
        ok = self.nranks <= 8
        if ok:
            try:
                ok = all(torch.cuda.can_device_access_peer(self.device, d) for d in range(torch.cuda.device_count())
                         if d != self.device) or torch.cuda.device_count() == 1
            except Exception:
                ok = False
        flags = [None] * self.nranks
        dist.all_gather_object(flags, bool(ok))
        if not all(flags):
            if self.rank == 0:
                warnings.warn("hdg_b200: peer-memory transport unavailable on some rank; using NCCL send/recv")
            return
        err = None
        try:
            self.p2p_setup()
        except Exception as exc:  # e.g. cudaIpcOpenMemHandle refused (MIG, containers without IPC)
            err = exc
        flags = [None] * self.nranks
        dist.all_gather_object(flags, err is None)
        if not all(flags):
            if self.p2p:
                self.p2p_enable(False)
                self.p2p = False
            if self.rank == 0:
                warnings.warn(f"hdg_b200: peer-memory transport could not be set up ({err}); using NCCL send/recv")

    def _collective_barrier(self):
        """ranks arrive at their first device-side exchange after rank-variable host work (scipy hierarchy / space
        construction): line them up so that no bounded spin of the peer-memory transport starts minutes early"""
        if self.part is None or self.part.nranks == 1:
            return
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                self.synchronize()
                dist.barrier()
        except ImportError:
            pass

    def p2p_setup(self, slab_doubles: int | None = None):
        """switch halo exchanges and all-reduces to the NVLink peer-memory transport: allocate this
        rank's mailbox, gather the CUDA IPC handles over torch.distributed, map the peers"""
        import torch.distributed as dist

        if slab_doubles is None:
            pt = self.part

            def biggest(plan):
                send = np.diff(plan.send_ptr).max(initial=0)
                return int(max(send, plan.recv_cnt.max(initial=0)))

            need = max(biggest(pt.cells) * 2 * self.nQ1, biggest(pt.facets) * (self.k + 2), biggest(pt.verts))
            box = [need]
            gathered = [None] * self.nranks
            dist.all_gather_object(gathered, box[0])
            slab_doubles = int(max(gathered)) + 1024  # the same on every rank
        buf = C.create_string_buffer(64)
        self._check(self.lib.hdg_p2p_alloc(self._h, int(slab_doubles), buf))
        handles = [None] * self.nranks
        dist.all_gather_object(handles, buf.raw)
        blob = C.create_string_buffer(b"".join(handles), 64 * self.nranks)
        self._check(self.lib.hdg_p2p_attach(self._h, blob))
        self.p2p = True
        dist.barrier()  # nobody pushes before everybody has mapped everybody

    def p2p_enable(self, on: bool = True):
        self._check(self.lib.hdg_p2p_enable(self._h, int(bool(on))))

    def p2p_status(self) -> int:
        err = C.c_int(0)
        self._check(self.lib.hdg_p2p_status(self._h, C.byref(err)))
        return err.value

    def halo_exchange_dev(self, kind: int, field, ndof: int):
        self._check(self.lib.hdg_halo_exchange_dev(self._h, int(kind), int(ndof), _dev(field)))

    def comm_stats(self):
        r, n, e, a = C.c_int(0), C.c_int(1), C.c_int64(0), C.c_int64(0)
        self._check(self.lib.hdg_comm_stats(self._h, C.byref(r), C.byref(n), C.byref(e), C.byref(a)))
        return {"rank": r.value, "nranks": n.value, "exchanges": e.value, "allreduces": a.value}

    # -- plumbing ---------------------------------------------------------------------------------
    def _check(self, rc, allow=()):
        if rc != HDG_OK and rc not in allow:
            raise HDGError(rc, self.lib.hdg_last_error(self._h).decode())
        return rc

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.hdg_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def use_torch_stream(self):
        import torch

        self._check(self.lib.hdg_set_stream(self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def synchronize(self):
        self._check(self.lib.hdg_synchronize(self._h))

    def shapes(self):
        return (self.nc, 2, self.nQ1), (self.nc, self.np_), (self.nf, self.nl1)

    # -- device buffers (torch owns the memory; SoA layout) ------------------------------------
    def empty(self, kind: int):
        import torch

        n = {0: 2 * self.nQ1 * self.nc, 1: self.np_ * self.nc, 2: self.nl1 * self.nf}[kind]
        return torch.empty(n, dtype=torch.float64, device=f"cuda:{self.device}")

    def zeros(self, kind: int):
        return self.empty(kind).zero_()

    def upload(self, kind: int, host_aos, out=None):
        a = _f64(host_aos)
        out = self.empty(kind) if out is None else out
        assert a.size == out.numel()
        self._check(self.lib.hdg_upload(self._h, kind, _ptr(a), _dev(out)))
        self.synchronize()  # `a` may be a temporary
        return out

    def download(self, kind: int, dev, out=None):
        shp = self.shapes()[kind]
        out = np.empty(shp, dtype=np.float64) if out is None else out
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == int(np.prod(shp))
        self._check(self.lib.hdg_download(self._h, kind, _dev(dev), _ptr(out)))
        return out

    # pipelined transfers on the engine's copy stream (host arrays must stay alive, and should be pinned, until copy_wait)
    def upload_begin(self, kind: int, host_aos, slot: int = 0):
        assert host_aos.dtype == np.float64 and host_aos.flags.c_contiguous
        self._check(self.lib.hdg_upload_begin(self._h, int(kind), _ptr(host_aos), int(slot)))

    def upload_end(self, kind: int, out, slot: int = 0):
        self._check(self.lib.hdg_upload_end(self._h, int(kind), int(slot), _dev(out)))
        return out

    def download_begin(self, kind: int, dev, out, slot: int = 0):
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == int(np.prod(self.shapes()[kind]))
        self._check(self.lib.hdg_download_begin(self._h, int(kind), _dev(dev), _ptr(out), int(slot)))

    def copy_wait(self):
        self._check(self.lib.hdg_copy_wait(self._h))

    def kernel_counts(self) -> dict:
        """launches per engine kernel since the engine was created (names without template arguments)"""
        need = self.lib.hdg_kernel_counts(self._h, None, 0)
        buf = C.create_string_buffer(int(need) + 64)
        self.lib.hdg_kernel_counts(self._h, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, _, cnt = line.partition("=")
            if name:
                out[name] = int(cnt)
        return out

    def kernel_times(self) -> dict:
        """in-situ device time per kernel since the last call: {name: (launches, milliseconds)}; needs
        set_tuning("ktime", 1) (diagnostics: event pairs around every launch, CUDA graphs off)"""
        need = self.lib.hdg_kernel_times(self._h, None, 0)
        buf = C.create_string_buffer(int(need) + 256)
        self.lib.hdg_kernel_times(self._h, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, _, rest = line.partition("=")
            cnt, _, ms = rest.partition(":")
            if name:
                out[name] = (int(cnt), float(ms))
        return out

    # -- condensed mixed-Poisson path ----------------------------------------------------------------
    def setup_poisson(self, keep_local: bool = False):
        self._check(self.lib.hdg_setup_poisson(self._h, int(keep_local)))

    def get_local_schur(self):
        """S_K as [nc, NL, NL]"""
        NL = 3 * self.nl1
        buf = np.empty((NL * NL, self.nc), dtype=np.float64)
        self._check(self.lib.hdg_get_local_schur(self._h, _ptr(buf)))
        return np.ascontiguousarray(buf.T).reshape(self.nc, NL, NL)

    def get_trace_matrix(self):
        """blocked ELL of P = -S: (val [nf,5,b,b], col [nf,5])"""
        b = self.nl1
        val = np.empty((self.nf, 5, b, b), dtype=np.float64)
        col = np.empty((self.nf, 5), dtype=np.int32)
        self._check(self.lib.hdg_get_trace_matrix(self._h, _ptr(val), col.ctypes.data_as(_ip)))
        return val, col

    def poisson_apply_host(self, rhs_Q=None, rhs_p=None, rhs_l=None, rtol=1e-12, maxit=10000, shift=True,
                           check=True):
        """solve the mixed-Poisson system for host residuals (AoS); returns (Q, p, l, iterations)"""
        rq, rp, rl = _f64(rhs_Q), _f64(rhs_p), _f64(rhs_l)
        sQ, sp_, sl = self.shapes()
        Q, p, l = np.empty(sQ), np.empty(sp_), np.empty(sl)
        its = C.c_int(0)
        rc = self.lib.hdg_poisson_apply_host(self._h, _ptr(rq), _ptr(rp), _ptr(rl), _ptr(Q), _ptr(p), _ptr(l),
                                             float(rtol), int(maxit), int(bool(shift)), C.byref(its))
        self._check(rc, allow=() if check else (HDG_ENOCONV,))
        self.last_iterations = its.value
        return Q, p, l, its.value

    def set_initial_guess(self, on: bool = True):
        """start the trace Krylov solve from the incoming trace vector instead of zero"""
        self._check(self.lib.hdg_set_initial_guess(self._h, int(bool(on))))

    def poisson_apply_dev(self, rhs_Q, rhs_p, rhs_l, Q, p, l, rtol=1e-12, maxit=10000, shift=True, check=True):
        its = C.c_int(0)
        rc = self.lib.hdg_poisson_apply_dev(self._h, _dev(rhs_Q), _dev(rhs_p), _dev(rhs_l), _dev(Q), _dev(p), _dev(l),
                                            float(rtol), int(maxit), int(bool(shift)), C.byref(its))
        self._check(rc, allow=() if check else (HDG_ENOCONV,))
        self.last_iterations = its.value
        return its.value

    def poisson_apply_update_dev(self, rhs_Q, rhs_p, rhs_l, Q_acc, p_acc, l, cq=0.0, cb=1.0, Q_base=None, cu=1.0,
                                 cp=0.0, rtol=1e-12, maxit=10000, check=True):
        """condensed solve whose back-substitution applies the caller's update in registers (k_back_update):
        Q_acc <- cq Q_acc + cb Q_base + cu u,  p_acc <- cp p_acc + phi,  l <- lam  with (u, phi, lam) the solution after
        the pressure shift; returns the trace-solve iteration count"""
        its = C.c_int(0)
        rc = self.lib.hdg_poisson_apply_update_dev(self._h, _dev(rhs_Q), _dev(rhs_p), _dev(rhs_l), float(cq), _dev(Q_acc),
                                                   float(cb), _dev(Q_base), float(cu), float(cp), _dev(p_acc), _dev(l),
                                                   float(rtol), int(maxit), C.byref(its))
        self._check(rc, allow=() if check else (HDG_ENOCONV,))
        self.last_iterations = its.value
        return its.value

    # -- multigrid preconditioner -------------------------------------------------------------------
    def mg_setup(self, hierarchy=None, smooth_fine=1, smooth_coarse=1, cheb_ratio=10.0, global_mesh=None,
                 repl_threshold=100_000):
        """build (host, scipy) and upload the GTMG hierarchy; the trace solve becomes MG-PCG.

        On a partitioned mesh `global_mesh` is the complete mesh: the hierarchy is built on it and
        split with :func:`partition.partition_hierarchy` (`hierarchy` may then be a ready
        :class:`partition.LocalHierarchy`)."""
        from . import multigrid

        dist = self.part is not None and self.part.nranks > 1
        if dist:
            from . import partition

            if hierarchy is None or not hasattr(hierarchy, "repl"):
                if global_mesh is None:
                    raise ValueError("mg_setup on a partitioned mesh needs global_mesh")
                Hg = multigrid.build_hierarchy(global_mesh, self.k) if hierarchy is None else hierarchy
                hierarchy = partition.partition_hierarchy(Hg, global_mesh, self.part, self.k,
                                                          repl_threshold=repl_threshold)
            for l, plan in enumerate(hierarchy.plans):
                self.set_halo_plan(2 + l, plan)
        H = multigrid.build_hierarchy(self.mesh, self.k) if hierarchy is None else hierarchy
        keep = []

        def conv(M):
            M = M.tocsr()
            M.sort_indices()
            rp = np.ascontiguousarray(M.indptr, dtype=np.int32)
            ci = np.ascontiguousarray(M.indices, dtype=np.int32)
            va = np.ascontiguousarray(M.data, dtype=np.float64)
            keep.extend([rp, ci, va])
            return hdg_csr(M.shape[0], M.shape[1], rp.ctypes.data_as(_ip), ci.ctypes.data_as(_ip), va.ctypes.data_as(_dp))

        nl = H.nlevels
        A = (hdg_csr * nl)(*[conv(a) for a in H.A])
        P = (hdg_csr * max(nl - 1, 1))(*[conv(p) for p in H.P]) if nl > 1 else None
        Rs = H.R if dist else [p.T for p in H.P]
        R = (hdg_csr * max(nl - 1, 1))(*[conv(r) for r in Rs]) if nl > 1 else None
        T = conv(H.T)
        Tt = conv(H.Tt if dist else H.T.T)
        lmax = np.ascontiguousarray(H.lmax, dtype=np.float64)
        pinv = np.ascontiguousarray(H.pinv, dtype=np.float64)
        self._collective_barrier()  # hdg_mg_setup runs power iterations with halo exchanges and all-reduces
        self._check(self.lib.hdg_mg_setup(self._h, nl, A, P, R, C.byref(T), C.byref(Tt), _ptr(lmax), _ptr(pinv),
                                          int(smooth_fine), int(smooth_coarse), float(cheb_ratio)))
        if dist:
            cnt = np.ascontiguousarray(H.gather_counts, dtype=np.int32)
            gid = np.ascontiguousarray(H.gather_gid, dtype=np.int32)
            self._check(self.lib.hdg_mg_set_distribution(self._h, int(H.repl), cnt.ctypes.data_as(_ip),
                                                         gid.ctypes.data_as(_ip)))
        self.hierarchy = H
        return H

    def mg_enable(self, on=True):
        self._check(self.lib.hdg_mg_enable(self._h, int(on)))

    def mg_info(self):
        nl, lm = C.c_int(0), C.c_double(0.0)
        self._check(self.lib.hdg_mg_info(self._h, C.byref(nl), C.byref(lm)))
        return nl.value, lm.value

    def trace_spmv_dev(self, x, y):
        self._check(self.lib.hdg_trace_spmv_dev(self._h, _dev(x), _dev(y)))

    def forward_eliminate_dev(self, rhs_Q, rhs_p, rhs_l, r_l):
        self._check(self.lib.hdg_forward_eliminate_dev(self._h, _dev(rhs_Q), _dev(rhs_p), _dev(rhs_l), _dev(r_l)))

    def back_substitute_dev(self, rhs_Q, rhs_p, l, Q, p):
        self._check(self.lib.hdg_back_substitute_dev(self._h, _dev(rhs_Q), _dev(rhs_p), _dev(l), _dev(Q), _dev(p)))

    # -- velocity side -----------------------------------------------------------------------------------
    def set_penalty(self, alpha: float):
        self._check(self.lib.hdg_set_penalty(self._h, float(alpha)))

    def set_tentative_solver(self, mode: int = 1, sweeps: int = 0):
        """0 = plain BiCGStab, 1 = facet-multiplier formulation with Chebyshev Schur sweeps (sweeps = 0 keeps the
        engine's current number, 4 by default)"""
        self._check(self.lib.hdg_set_tentative_solver(self._h, int(mode), int(sweeps)))

    def tentative_stats(self):
        """counters of the tentative-velocity solver since the engine was created (`hdg_tentative_stats`)"""
        out = (C.c_int64 * 6)()
        self._check(self.lib.hdg_tentative_stats(self._h, out))
        keys = ("solves", "bicgstab_iterations", "fgmres_iterations", "fallbacks", "failed_verifications", "fgmres_cycles")
        return dict(zip(keys, [int(v) for v in out]))

    def comm_probe(self, kind: int, ndof: int, nred: int = 2, nrep: int = 200):
        """(microseconds per halo exchange, per all-reduce) in isolation; collective (`hdg_comm_probe`)"""
        a, b = C.c_double(0.0), C.c_double(0.0)
        self._check(self.lib.hdg_comm_probe(self._h, int(kind), int(ndof), int(nred), int(nrep), C.byref(a), C.byref(b)))
        return a.value, b.value

    def mixed_stats(self):
        """counters of the mixed-precision tentative solver (`hdg_mixed_stats`)"""
        out = (C.c_int64 * 4)()
        self._check(self.lib.hdg_mixed_stats(self._h, out))
        return dict(zip(("solves", "outer_steps", "inner_fp32_iterations", "handed_to_fp64"), [int(v) for v in out]))

    def set_tentative_comm(self, local_sweeps: bool = False):
        """multi-GPU: skip (True) or perform (False) the halo exchanges between Schur sweeps"""
        self._check(self.lib.hdg_set_tentative_comm(self._h, int(bool(local_sweeps))))

    def project_bdm_dev(self, Q, Qstar):
        self._check(self.lib.hdg_project_bdm_dev(self._h, _dev(Q), _dev(Qstar)))

    def fimpl_apply_dev(self, Qstar, X, Y, c0=0.0, c1=1.0, upwind=True):
        self._check(self.lib.hdg_fimpl_apply_dev(self._h, _dev(Qstar), _dev(X), float(c0), float(c1), int(upwind), _dev(Y)))

    def tentative_solve_dev(self, Qstar, adt, rhs, x, upwind=True, rtol=1e-10, maxit=20000, zero_guess=True,
                            check=True):
        its = C.c_int(0)
        rc = self.lib.hdg_tentative_solve_dev(self._h, _dev(Qstar), float(adt), int(upwind), _dev(rhs), _dev(x),
                                              float(rtol), int(maxit), int(zero_guess), C.byref(its))
        self._check(rc, allow=() if check else (HDG_ENOCONV,))
        return its.value

    def weak_divergence_dev(self, Q, Rp, scale=1.0, mode=1):
        self._check(self.lib.hdg_weak_divergence_dev(self._h, _dev(Q), float(scale), int(mode), _dev(Rp)))

    def pressure_gradient_dev(self, p, l, Y, c0=0.0, c1=1.0):
        self._check(self.lib.hdg_pressure_gradient_dev(self._h, _dev(p), _dev(l), float(c0), float(c1), _dev(Y)))

    def reconstruct_trace_dev(self, Q, p, l):
        self._check(self.lib.hdg_reconstruct_trace_dev(self._h, _dev(Q), _dev(p), _dev(l)))

    def shift_pressure_dev(self, p, l=None):
        self._check(self.lib.hdg_shift_pressure_dev(self._h, _dev(p), _dev(l)))

    def reconstruction_rhs_dev(self, Q, b, Rp, Rl):
        self._check(self.lib.hdg_reconstruction_rhs_dev(self._h, _dev(Q), _dev(b), _dev(Rp), _dev(Rl)))

    def gamma_apply_dev(self, Q, p, l, Rp, Rl):
        """(Rp, Rl) = Gamma(psi, mu; Q, p, l), the constraint rows of the monolithic operator"""
        self._check(self.lib.hdg_gamma_apply_dev(self._h, _dev(Q), _dev(p), _dev(l), _dev(Rp), _dev(Rl)))

    def dot_dev(self, kind, x, y):
        out = C.c_double(0.0)
        self._check(self.lib.hdg_dot_dev(self._h, int(kind), _dev(x), _dev(y), C.byref(out)))
        return out.value

    def l2_inner_dev(self, kind, x, y):
        out = C.c_double(0.0)
        self._check(self.lib.hdg_l2_inner_dev(self._h, int(kind), _dev(x), _dev(y), C.byref(out)))
        return out.value

    def lincomb_dev(self, out, terms):
        """out = sum c_t * x_t for terms = [(c, tensor), ...] (at most 8; out may alias an input)"""
        n = len(terms)
        coefs = (C.c_double * n)(*[float(c) for c, _ in terms])
        ptrs = (_vp * n)(*[t.data_ptr() for _, t in terms])
        for _, t in terms:
            assert t.numel() == out.numel()
        self._check(self.lib.hdg_lincomb_dev(self._h, out.numel(), _dev(out), n, coefs, ptrs))

    def mass_dev(self, kind, x, y, inverse=False):
        self._check(self.lib.hdg_mass_dev(self._h, int(kind), int(inverse), _dev(x), _dev(y)))

    # -- passive tracer ----------------------------------------------------------
    PLAN_CG = 18  # halo-plan kind of the CG dofs (include/hdg_b200.h)

    def tracer_setup(self, nq_facet: int | None = None, global_mesh=None):
        """build the [CG_{k+1}]^2 space of the velocity projection (`common.py:119-122`) and hand it and
        the quadrature tables of the advection kernel to the engine; idempotent.  On a partitioned mesh
        `global_mesh` is the complete mesh: the dofs are renumbered owned-first and their halo plan is
        installed (``partition.cg_plan``)."""
        if getattr(self, "cg_space", None) is not None and getattr(self, "_tracer_nq_facet", None) == nq_facet:
            return self.cg_space
        from . import cgspace

        n_owned = None
        if self.part is not None and self.part.nranks > 1:
            if global_mesh is None:
                raise ValueError("tracer_setup on a partitioned mesh needs global_mesh")
            from . import partition

            plan, perm = partition.cg_plan(global_mesh, self.part, self.k + 1)
            sp = cgspace.build_cg_space(self.mesh, self.k + 1, perm=perm)
            self.set_halo_plan(self.PLAN_CG, plan)
            n_owned = plan.n_owned
        else:
            sp = cgspace.build_cg_space(self.mesh, self.k + 1)
        tab_cell, tab_facet = cgspace.tracer_tables(self.k, nq_facet)
        cellmap = np.ascontiguousarray(sp.cellmap.T, dtype=np.int32)  # SoA [nloc][nc]
        dinv = np.ascontiguousarray(1.0 / sp.diag)
        self._collective_barrier()
        self._check(self.lib.hdg_tracer_setup(
            self._h, sp.ndof, sp.ndof if n_owned is None else n_owned, cellmap.ctypes.data_as(_ip),
            sp.inc_ptr.ctypes.data_as(_ip), sp.inc_idx.ctypes.data_as(_ip), _ptr(sp.W), _ptr(dinv), tab_cell.shape[0],
            _ptr(tab_cell), tab_facet.shape[1], _ptr(tab_facet)))
        self.cg_space = sp
        self._tracer_nq_facet = nq_facet
        return sp

    def project_cg_dev(self, Q, Qcg, rtol=1e-13, maxit=500):
        """Qcg = cell-wise representation of the L2 projection of Q onto [CG_{k+1}]^2; returns the
        number of PCG iterations"""
        its = C.c_int(0)
        self._check(self.lib.hdg_project_cg_dev(self._h, _dev(Q), _dev(Qcg), float(rtol), int(maxit), C.byref(its)))
        return its.value

    def tracer_advection_dev(self, Qcg, q, out, c0=0.0, acc=None, c1=1.0):
        """out = c0 acc + c1 M^-1 _tracer_advection(chi, q, Qcg)  (`common.py:110-129`)"""
        self._check(self.lib.hdg_tracer_advection_dev(self._h, _dev(Qcg), _dev(q), float(c0),
                                                      _dev(acc) if acc is not None else None, float(c1), _dev(out)))

    # -- reporting -----------------------------------------------------------------------------------
    def timers(self):
        n = len(TIMER_LABELS)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        self._check(self.lib.hdg_get_timers(self._h, ms, cnt, n))
        return {lab: (ms[i], cnt[i]) for i, lab in enumerate(TIMER_LABELS)}

    def measure_fp64_peak(self) -> float:
        """measured FP64 FMA throughput in TFLOP/s"""
        out = C.c_double(0.0)
        self._check(self.lib.hdg_measure_fp64_peak(self._h, C.byref(out)))
        return out.value

    def tent_sweep_probe(self, adt: float, nrep: int = 20) -> float:
        """mean ms of `nrep` back-to-back k_tent_sweep launches (facet Schur sweep of the tentative solver)"""
        out = C.c_double(0.0)
        self._check(self.lib.hdg_tent_sweep_probe(self._h, float(adt), int(nrep), C.byref(out)))
        return out.value

    def set_tuning(self, name: str, value: int):
        """result-neutral tuning knobs, e.g. ("sweep_minblocks", 6)"""
        self._check(self.lib.hdg_set_tuning(self._h, name.encode(), int(value)))

    def debug_scalars(self):
        out = (C.c_double * 10)()
        self._check(self.lib.hdg_debug_scalars(self._h, out))
        v = list(out)
        return {"cg": {"ref": v[0], "rz": v[1], "tol2": v[2], "iters": int(v[3]), "done": int(v[4])},
                "bicgstab": {"bb": v[5], "rr": v[6], "tol2": v[7], "iters": int(v[8]), "done": int(v[9])}}

    def set_graphs(self, on: bool = True):
        """replay the Krylov iteration bodies as CUDA graphs (default) or launch kernel by kernel"""
        self._check(self.lib.hdg_set_graphs(self._h, int(bool(on))))

    @property
    def graph_replays(self) -> int:
        n = C.c_int64(0)
        self._check(self.lib.hdg_graph_replays(self._h, C.byref(n)))
        return n.value

    @property
    def guess_restarts(self) -> int:
        n = C.c_int64(0)
        self._check(self.lib.hdg_guess_restarts(self._h, C.byref(n)))
        return n.value

    def reset_timers(self):
        self._check(self.lib.hdg_reset_timers(self._h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.hdg_launch_count(self._h))
