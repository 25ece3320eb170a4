"""Tensor/array-level API over the C ABI: build HEM pyramids on a B200.

``HemContext.build`` accepts numpy arrays (host memory; copied H2D inside the
call) or torch CUDA tensors (used in place through their data pointers) and
returns per-level arrays in the reference's attribute layout
(reference src/cpp_ext/src/mixturelevel.cpp:30-70).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi

STAT_NAMES = ("n_in", "n_parents", "n_pairs", "n_orphans", "n_dropped", "n_out", "grid_cells", "cell_size",
              "max_radius", "ms_total", "ms_prep", "ms_grid", "ms_gate", "ms_sumlw", "ms_moments", "ms_compact", "ms_k_gate", "ms_k_moments", "n_certified_parents", "ms_host_enqueue",
              "n_parents_local", "n_pairs_local", "shards", "tight_box")


@dataclass
class HemParams:
    """Mirror of the reference's GaussianMixtureParams (src/params/merge_parameters.py:4-10)."""
    hem_reduction: float = 3.0
    distance_delta: float = 3.0
    color_delta: float = 2.5
    decay_rate: float = 1.0
    cluster_level: int = 3


@dataclass
class IcpResult:
    """Field names of Open3D's RegistrationResult (what the reference reads:
    qt_multiscale_registrator.py:151-161) plus diagnostics."""
    transformation: np.ndarray
    fitness: float
    inlier_rmse: float
    iterations: int
    n_correspondences: int
    stop_reason: int          # 1 converged, 2 max_iteration, 3 fewer than 3 correspondences
    ms_device: float
    grid_cells: int
    cell_size: float
    correspondence: np.ndarray | None = None     # [n_source] target index or -1

    @property
    def correspondence_set(self) -> np.ndarray:
        """[K,2] (source index, target index) pairs like Open3D's correspondence_set."""
        if self.correspondence is None:
            raise ValueError("run with with_correspondences=True")
        src = np.nonzero(self.correspondence >= 0)[0]
        return np.stack([src, self.correspondence[src]], 1).astype(np.int32)

    @staticmethod
    def from_flat(r, corr) -> "IcpResult":
        return IcpResult(np.array(r[:16]).reshape(4, 4), float(r[16]), float(r[17]), int(r[18]), int(r[19]), int(r[20]),
                         float(r[21]), int(r[22]), float(r[23]), corr)


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x, device=None, dtype="torch.float32"):
    """(pointer, keepalive) of a numpy array or torch tensor, None passes through.  A torch CUDA tensor is used in
    place through its data pointer, so it must already have the element type the C ABI reads and live on the
    context's GPU: anything else would be reinterpreted silently."""
    if x is None:
        return None, None
    if _is_torch(x):
        if x.is_cuda:
            if device is not None and x.device.index != device:
                raise ValueError(f"tensor is on cuda:{x.device.index}, the context runs on cuda:{device}")
            if dtype is not None and str(x.dtype) != dtype:
                raise TypeError(f"expected a {dtype} tensor, got {x.dtype}")
        elif dtype is not None and str(x.dtype) != dtype:
            x = x.to(getattr(__import__("torch"), dtype.split(".")[1]))
        if not x.is_contiguous():
            x = x.contiguous()
        return C.c_void_p(x.data_ptr()), x
    a = np.ascontiguousarray(x)
    return C.c_void_p(a.ctypes.data), a


def slab_range(n_parents: int, rank: int, world: int) -> tuple[int, int]:
    """Parents (positions in grid-cell order) [begin, end) owned by `rank` — host arithmetic only."""
    lib = _capi.load()
    b, e = C.c_int64(0), C.c_int64(0)
    rc = lib.hem_slab_range(int(n_parents), int(rank), int(world), C.byref(b), C.byref(e))
    if rc != 0:
        raise _capi.HemError(rc, "bad slab_range arguments")
    return int(b.value), int(e.value)


def nccl_unique_id() -> bytes:
    lib = _capi.load()
    buf = (C.c_uint8 * _capi.NCCL_ID_BYTES)()
    rc = lib.hem_nccl_unique_id(buf)
    if rc != 0:
        msg = lib.hem_last_error(None)
        raise _capi.HemError(rc, msg.decode() if msg else "hem_nccl_unique_id failed")
    return bytes(buf)


def exchange_unique_id(rank=None, world=None, make_id=nccl_unique_id):
    """Rank 0 creates the id, everyone receives it through torch.distributed (gloo or nccl)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.zeros(_capi.NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(make_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes()), rank, world


class HemContext:
    """One context per GPU (``hem_create`` / ``hem_destroy``)."""

    def __init__(self, device: int = 0, seed: int = 0):
        self._lib = _capi.load()
        self._h = self._lib.hem_create(int(device), int(seed))
        if not self._h:
            msg = self._lib.hem_last_error(None)
            raise _capi.HemError(-3, msg.decode() if msg else "hem_create failed")
        self.device = device
        self.sh_floats = 0
        self.rank, self.world = 0, 1
        self._user_stream = False          # set_stream() was called: the caller owns the ordering
        self._borrowed = None              # device inputs the library aliases beyond the call (level-0 SH rows)

    # -- plumbing -----------------------------------------------------------
    def _check(self, code: int):
        if code != 0:
            msg = self._lib.hem_last_error(self._h)
            raise _capi.HemError(code, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, name: str, value: float):
        """hem_set_option (include/hem_b200.h lists every option): e.g. "timing" (1: events around every stage, 2: level
        and hot kernels only), "keep_pairs", "geo_first" (host SH input: geometry rows first, SH passes once the upload has
        landed), "overlap", "poll_sync", "tight_box", "heavy_rows", "shard_min_n".  Results never depend on them."""
        self._check(self._lib.hem_set_option(self._h, name.encode(), float(value)))

    def set_stream(self, stream_handle: int | None):
        self._user_stream = stream_handle is not None
        self._check(self._lib.hem_set_stream(self._h, C.c_void_p(stream_handle or 0)))

    def _follow_torch_stream(self, *arrays):
        """Device tensors are produced and consumed asynchronously on torch's current stream: unless the caller
        chose a stream, run on that one so the library's reads and writes are ordered with the caller's kernels."""
        if self._user_stream:
            return
        if any(a is not None and _is_torch(a) and a.is_cuda for a in arrays):
            self._check(self._lib.hem_set_stream(self._h, C.c_void_p(self._torch_stream())))

    def _torch_stream(self) -> int:
        """Handle of torch's current stream on the context's GPU; the legacy default stream (handle 0 in torch) is
        passed as cudaStreamLegacy, because 0 means "the context's own stream" to hem_set_stream."""
        import torch
        return torch.cuda.current_stream(self.device).cuda_stream or 1

    # -- multi-GPU ------------------------------------------------------------------
    def init_distributed(self, rank: int | None = None, world: int | None = None, unique_id: bytes | None = None):
        """Join the NCCL communicator of a one-process-per-GPU job.  With
        no arguments the rank/world come from ``torch.distributed`` (already
        initialised, any backend) and rank 0's NCCL unique id is broadcast through it."""
        if unique_id is None:
            unique_id, rank, world = exchange_unique_id(rank, world)
        buf = (C.c_uint8 * _capi.NCCL_ID_BYTES).from_buffer_copy(unique_id)
        self._check(self._lib.hem_comm_init(self._h, buf, int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    # -- the hot path ----------------------------------------------------------
    def build(self, xyz, colors, opacity, cov6, sh, cluster_level: int, hem_reduction: float = 3.0,
              distance_delta: float = 3.0, color_delta: float = 2.5, decay_rate: float = 1.0,
              weight=None, parent_mask=None) -> int:
        """Run ``cluster_level`` HEM steps (reference MixtureCreator.CreateMixture,
        src/cpp_ext/mixture_wrapper.cpp:10-18).  Returns the number of levels held
        (cluster_level + 1, level 0 = input)."""
        n = int(xyz.shape[0])
        F = int(sh.shape[1]) if sh is not None and len(sh.shape) == 2 else 0
        if not _is_torch(xyz):
            xyz, colors, opacity, cov6 = (np.ascontiguousarray(a, dtype=np.float32) for a in (xyz, colors, opacity, cov6))
            sh = None if F == 0 else np.ascontiguousarray(sh, dtype=np.float32)
            weight = None if weight is None else np.ascontiguousarray(weight, dtype=np.float32)
            parent_mask = None if parent_mask is None else np.ascontiguousarray(parent_mask, dtype=np.uint8)
        self._follow_torch_stream(xyz, colors, opacity, cov6, sh, weight, parent_mask)
        keep = []
        ptrs = []
        for a in (xyz, colors, opacity, cov6, sh if F else None, weight):
            p, k = _ptr(a, self.device)
            ptrs.append(p)
            keep.append(k)
        pm, k = _ptr(parent_mask, self.device, "torch.uint8")
        keep.append(k)
        # a device-resident SH array is used in place as level 0 (no copy): it has to outlive the call for
        # level(0) / level_torch(0) to stay valid, so the context holds on to it until the next build
        self._borrowed = keep[4] if F and _is_torch(keep[4]) and keep[4].is_cuda else None
        self.sh_floats = F
        self._check(self._lib.hem_build_pyramid(self._h, *ptrs, n, F, int(cluster_level), float(hem_reduction),
                                                float(distance_delta), float(color_delta), float(decay_rate), pm))
        return self.num_levels

    def registration_inputs(self, level: int, normals: bool = True) -> dict:
        """What the reference hands to Open3D for a level (convert_gs_to_open3d_pc,
        src/utils/point_cloud_converter.py:31-49), computed on the GPU: float64 points, sh2rgb
        colours, [N,3,3] covariances and (optionally) covariance-derived normals."""
        n = self.level_size(level)
        out = dict(points=np.empty((n, 3), np.float64), colors=np.empty((n, 3), np.float64),
                   covariances=np.empty((n, 3, 3), np.float64))
        if normals:
            out["normals"] = np.empty((n, 3), np.float64)
        self._check(self._lib.hem_level_registration_inputs(
            self._h, level, C.c_void_p(out["points"].ctypes.data), C.c_void_p(out["colors"].ctypes.data),
            C.c_void_p(out["covariances"].ctypes.data),
            C.c_void_p(out["normals"].ctypes.data) if normals else None))
        return out

    def from_mixture(self, level: int, sh_degree: int | None = None, decomposition: bool = False) -> dict:
        """The attributes GaussianModel.from_mixture builds from a mixture level (reference
        src/models/gaussian_model.py:141-153), as torch CUDA tensors with the reference's names and shapes:
        `_xyz` [n,3], `_features_dc` [n,1,3], `_features_rest` [n,(deg+1)^2-1,3], `_opacity` [n] (1-D, as in the
        reference), `_covariance` [n,6], `_scaling` [n,3] and `_rotation` [n,4] from the GPU eigen-decomposition
        (``hem_level_from_mixture``; the reference runs torch.linalg.eigh + Python scatter code here)."""
        import torch
        d = self.level_torch(level)
        n, F = d["xyz"].shape[0], self.sh_floats
        if sh_degree is None:
            sh_degree = int(round((F // 3 + 1) ** 0.5)) - 1
        if 3 * ((sh_degree + 1) ** 2 - 1) != F:
            raise ValueError(f"sh_degree {sh_degree} does not match {F} SH floats per splat")
        dev = d["xyz"].device
        scaling = torch.empty((n, 3), dtype=torch.float32, device=dev)
        rotation = torch.empty((n, 4), dtype=torch.float32, device=dev)
        ev = torch.empty((n, 3), dtype=torch.float32, device=dev) if decomposition else None
        evec = torch.empty((n, 3, 3), dtype=torch.float32, device=dev) if decomposition else None
        self._check(self._lib.hem_level_from_mixture(
            self._h, level, C.c_void_p(scaling.data_ptr()), C.c_void_p(rotation.data_ptr()),
            C.c_void_p(ev.data_ptr()) if decomposition else None, C.c_void_p(evec.data_ptr()) if decomposition else None))
        out = dict(sh_degree=sh_degree, _xyz=d["xyz"], _features_dc=d["colors"].view(-1, 1, 3),
                   _features_rest=d["sh"].view(n, (sh_degree + 1) ** 2 - 1, 3), _opacity=d["opacity"],
                   _covariance=d["cov6"], _scaling=scaling, _rotation=rotation)
        if decomposition:
            out["eigenvalues"], out["eigenvectors"] = ev, evec
        return out

    def icp(self, source, target, init=None, max_correspondence: float = 5.0, relative_fitness: float = 1e-6,
            relative_rmse: float = 1e-6, max_iteration: int = 30, with_correspondences: bool = False) -> "IcpResult":
        """Point-to-point ICP on the GPU with Open3D's definitions (``hem_icp_point_to_point``; replaces
        the registration_icp call of the reference's do_icp_registration,
        src/utils/local_registration_util.py:76-100).  source/target: [n,3] float64 numpy arrays or
        torch CUDA tensors."""
        self._follow_torch_stream(source, target)
        ps, ks = _ptr(source if _is_torch(source) else np.asarray(source, np.float64), self.device, "torch.float64")
        pt, kt = _ptr(target if _is_torch(target) else np.asarray(target, np.float64), self.device, "torch.float64")
        ns, nt = int(ks.shape[0]), int(kt.shape[0])
        res = np.zeros(_capi.ICP_RESULT_LEN, np.float64)
        T0 = None if init is None else np.ascontiguousarray(np.asarray(init, np.float64).reshape(4, 4))
        corr = np.empty(ns, np.int32) if with_correspondences else None
        self._check(self._lib.hem_icp_point_to_point(
            self._h, ps, ns, pt, nt, None if T0 is None else C.c_void_p(T0.ctypes.data), float(max_correspondence),
            float(relative_fitness), float(relative_rmse), int(max_iteration),
            res.ctypes.data_as(C.POINTER(C.c_double)), None if corr is None else C.c_void_p(corr.ctypes.data)))
        return IcpResult.from_flat(res, corr)

    def icp_multiscale(self, target_ctx: "HemContext", max_correspondences, max_iterations, init=None,
                       relative_fitness: float = 1e-6, relative_rmse: float = 1e-6) -> list:
        """Coarse-to-fine ICP of this context's pyramid (source) onto ``target_ctx``'s pyramid
        (``hem_icp_multiscale``; the loop of MultiScaleRegistratorMixture._register_main_point_clouds,
        src/gui/workers/registration/qt_multiscale_registrator.py:203-230): scale i uses level
        ``num_levels-1-i`` of both pyramids.  Returns one IcpResult per scale; the last one is the answer."""
        k = len(max_correspondences)
        if k != len(max_iterations):
            raise ValueError("The number of iteration and voxel values provided do not match.")
        mc = np.ascontiguousarray(max_correspondences, np.float64)
        mi = np.ascontiguousarray(max_iterations, np.int32)
        T0 = None if init is None else np.ascontiguousarray(np.asarray(init, np.float64).reshape(4, 4))
        res = np.zeros((k, _capi.ICP_RESULT_LEN), np.float64)
        self._check(self._lib.hem_icp_multiscale(
            self._h, target_ctx._h, k, mc.ctypes.data_as(C.POINTER(C.c_double)), mi.ctypes.data_as(C.POINTER(C.c_int32)),
            float(relative_fitness), float(relative_rmse), None if T0 is None else C.c_void_p(T0.ctypes.data),
            res.ctypes.data_as(C.POINTER(C.c_double))))
        return [IcpResult.from_flat(r, None) for r in res]

    def bind_output(self, level: int, xyz=None, colors=None, opacity=None, cov6=None, sh=None):
        """Bind caller-owned destination buffers (numpy views of pinned memory, or torch
        tensors) for `level`; the next ``build`` copies the level out on a second stream as
        soon as it is complete (``hem_bind_output``).  Row capacity = rows of the smallest buffer."""
        bufs = [b for b in (xyz, colors, opacity, cov6, sh) if b is not None]
        cap = min(int(b.shape[0]) for b in bufs) if bufs else 0
        ptrs = []
        for a in (xyz, colors, opacity, cov6, sh):
            if a is None:
                ptrs.append(None)
            elif _is_torch(a):
                ptrs.append(C.c_void_p(a.data_ptr()))
            else:
                assert a.flags["C_CONTIGUOUS"] and a.dtype == np.float32
                ptrs.append(C.c_void_p(a.ctypes.data))
        self._bound = getattr(self, "_bound", {})
        self._bound[level] = (xyz, colors, opacity, cov6, sh)          # keep alive
        self._check(self._lib.hem_bind_output(self._h, int(level), *ptrs, cap))

    def clear_outputs(self):
        self._bound = {}
        self._check(self._lib.hem_clear_outputs(self._h))

    @property
    def num_levels(self) -> int:
        return int(self._lib.hem_num_levels(self._h))

    def level_size(self, level: int) -> int:
        n = int(self._lib.hem_level_size(self._h, level))
        if n < 0:
            raise _capi.HemError(-4, f"level {level} not built")
        return n

    def level(self, level: int, with_internal: bool = False) -> dict:
        """One level as host numpy arrays (reference CreatePythonLists layout)."""
        n, F = self.level_size(level), self.sh_floats
        out = dict(xyz=np.empty((n, 3), np.float32), colors=np.empty((n, 3), np.float32),
                   opacity=np.empty(n, np.float32), cov6=np.empty((n, 6), np.float32),
                   sh=np.empty((n, F), np.float32))
        w = f = None
        if with_internal:
            out["weight"] = np.empty(n, np.float32)
            out["is_parent"] = np.empty(n, np.uint8)
            w, f = C.c_void_p(out["weight"].ctypes.data), C.c_void_p(out["is_parent"].ctypes.data)
        self._check(self._lib.hem_level_copy(
            self._h, level, C.c_void_p(out["xyz"].ctypes.data), C.c_void_p(out["colors"].ctypes.data),
            C.c_void_p(out["opacity"].ctypes.data), C.c_void_p(out["cov6"].ctypes.data),
            C.c_void_p(out["sh"].ctypes.data) if F else None, w, f))
        return out

    def level_into(self, level: int, xyz=None, colors=None, opacity=None, cov6=None, sh=None,
                   weight=None, is_parent=None) -> int:
        """Copy one level into caller-owned buffers (numpy arrays, e.g. views of pinned
        torch tensors, or torch CUDA tensors); returns the level size.  Buffers must be
        contiguous float32 (uint8 for is_parent) with room for level_size rows."""
        ptrs = []
        keep = []
        for a in (xyz, colors, opacity, cov6, sh if self.sh_floats else None, weight, is_parent):
            if a is None:
                ptrs.append(None)
            elif _is_torch(a):
                ptrs.append(C.c_void_p(a.data_ptr()))
            else:
                assert a.flags["C_CONTIGUOUS"]
                ptrs.append(C.c_void_p(a.ctypes.data))
            keep.append(a)
        self._check(self._lib.hem_level_copy(self._h, level, *ptrs))
        return self.level_size(level)

    def level_torch(self, level: int, with_internal: bool = False) -> dict:
        """One level as torch CUDA tensors (device-to-device, no host staging)."""
        import torch
        n, F = self.level_size(level), self.sh_floats
        dev = torch.device("cuda", self.device)
        if not self._user_stream:       # the new tensors belong to torch's current stream: write them on it
            self._check(self._lib.hem_set_stream(self._h, C.c_void_p(self._torch_stream())))
        out = dict(xyz=torch.empty((n, 3), dtype=torch.float32, device=dev),
                   colors=torch.empty((n, 3), dtype=torch.float32, device=dev),
                   opacity=torch.empty(n, dtype=torch.float32, device=dev),
                   cov6=torch.empty((n, 6), dtype=torch.float32, device=dev),
                   sh=torch.empty((n, F), dtype=torch.float32, device=dev))
        w = f = None
        if with_internal:
            out["weight"] = torch.empty(n, dtype=torch.float32, device=dev)
            out["is_parent"] = torch.empty(n, dtype=torch.uint8, device=dev)
            w, f = C.c_void_p(out["weight"].data_ptr()), C.c_void_p(out["is_parent"].data_ptr())
        self._check(self._lib.hem_level_copy(
            self._h, level, C.c_void_p(out["xyz"].data_ptr()), C.c_void_p(out["colors"].data_ptr()),
            C.c_void_p(out["opacity"].data_ptr()), C.c_void_p(out["cov6"].data_ptr()),
            C.c_void_p(out["sh"].data_ptr()) if F and n else None, w, f))
        return out

    def step_stats(self, step: int) -> dict:
        buf = (C.c_double * _capi.STEP_STATS_LEN)()
        self._check(self._lib.hem_step_stats(self._h, step, buf))
        d = dict(zip(STAT_NAMES, list(buf)))
        for k in STAT_NAMES[:7] + STAT_NAMES[18:19] + STAT_NAMES[20:24]:
            d[k] = int(d[k])
        return d

    def step_stats_table(self, steps: int):
        """hem_step_stats of steps 0..steps-1 as one float64 array [steps, len(STAT_NAMES)] (columns in STAT_NAMES order):
        the cheap form for callers that read the statistics after every build."""
        import numpy as np
        out = np.empty((steps, _capi.STEP_STATS_LEN), dtype=np.float64)
        row = C.POINTER(C.c_double)
        base = out.ctypes.data
        for l in range(steps):
            self._check(self._lib.hem_step_stats(self._h, l, C.cast(base + l * out.strides[0], row)))
        return out

    def info(self, name: str) -> float:
        out = C.c_double(0)
        self._check(self._lib.hem_get_info(self._h, name.encode(), C.byref(out)))
        return float(out.value)

    def step_pairs(self, step: int) -> dict:
        """White-box child lists of a step (needs option keep_pairs=1 before build)."""
        s = self.step_stats(step)
        npar, P, n = s["n_parents"], s["n_pairs"], s["n_in"]
        out = dict(parent_idx=np.zeros(npar, np.uint32), off=np.zeros(npar + 1, np.int64),
                   child=np.zeros(P, np.uint32), wl=np.zeros(P, np.float32), sumlw=np.zeros(n, np.float32),
                   radius=np.zeros(npar, np.float32))
        self._check(self._lib.hem_step_pairs(
            self._h, step, out["parent_idx"].ctypes.data_as(_capi._u32), out["off"].ctypes.data_as(_capi._i64),
            out["child"].ctypes.data_as(_capi._u32), out["wl"].ctypes.data_as(_capi._f),
            out["sumlw"].ctypes.data_as(_capi._f), out["radius"].ctypes.data_as(_capi._f)))
        return out
