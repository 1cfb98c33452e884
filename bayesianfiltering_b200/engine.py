"""Thin owners of libbfilt handles.  Two calling modes per engine:

  * device mode: torch CUDA tensors in, torch CUDA tensors out, work enqueued on torch's current stream,
    no synchronisation (`run`);
  * host mode: NumPy arrays in, NumPy arrays out through the `*_run_host` C entry points, which do the
    H2D / D2H copies themselves (`run_host`) -- this is the end-to-end path bench.py times.

Buffers are time-major ([B,T,N,..] / [B,T,M,..]); `inference.py` returns the reference's [N,T,..] views.
"""
import ctypes as C

import numpy as np

from . import _lib

_RESAMPLERS = {"multinomial": _lib.RESAMPLE_MULTINOMIAL, "systematic": _lib.RESAMPLE_SYSTEMATIC}
_PATHS = {"auto": _lib.PATH_AUTO, "resident": _lib.PATH_RESIDENT, "streamed": _lib.PATH_STREAMED}

PF_OUTPUTS = {
    # name: (io field, per-(trial,step) shape as a function of (N, dx), dtype)
    "particles": ("out_particles", lambda N, dx: (N, dx), np.float32),
    "weights": ("out_weights", lambda N, dx: (N,), np.float32),
    "propagated": ("out_propagated", lambda N, dx: (N, dx), np.float32),
    "norm_weights": ("out_norm_weights", lambda N, dx: (N,), np.float32),
    "ancestors": ("out_ancestors", lambda N, dx: (N,), np.int32),
    "means": ("out_mean", lambda N, dx: (dx,), np.float32),
    "log_likelihood": ("out_loglik", lambda N, dx: (), np.float32),
    "ess": ("out_ess", lambda N, dx: (), np.float32),
    "resampled": ("out_resampled", lambda N, dx: (), np.int32),
}
PF_FINALS = {
    "final_particles": ("final_particles", lambda N, dx: (N, dx)),
    "final_weights": ("final_weights", lambda N, dx: (N,)),
}
GSF_OUTPUTS = {
    "weights": ("out_weights", lambda M, dx: (M,)),
    "means": ("out_means", lambda M, dx: (M, dx)),
    "covariances": ("out_covs", lambda M, dx: (M, dx, dx)),
    "predicted_means": ("out_pred_means", lambda M, dx: (M, dx)),
    "predicted_covariances": ("out_pred_covs", lambda M, dx: (M, dx, dx)),
    "log_likelihood": ("out_loglik", lambda M, dx: ()),
    "mixture_mean": ("out_mix_mean", lambda M, dx: (dx,)),
}


def _default_device():
    import os
    return int(os.environ.get("LOCAL_RANK", "0"))


def _np32(a, shape, name):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if a.size != int(np.prod(shape)):
        raise ValueError(f"{name} has {a.size} elements, expected shape {tuple(shape)}")
    return a.reshape(shape)


def _host_empty(shape, dtype):
    """Result array in page-locked host memory when it is big enough to matter: the device-to-host copy inside
    `bf_*_run_host` then runs at PCIe speed instead of through the driver's pageable staging (10x for the 1 GB the
    Gaussian-sum filter returns at M=100 x 1024 trials x T=500).  torch's pinned allocator caches the pages."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if n >= (1 << 20):
        try:
            import torch
            tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
            return torch.empty(tuple(shape), dtype=tdt, pin_memory=True).numpy()
        except Exception:
            pass
    return np.empty(shape, dtype=dtype)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


class ParticleFilterEngine:
    """Owns a bf_pf_handle: N particles x B independent trials of one registry model."""

    def __init__(self, host_model, num_particles, num_trials=1, ess_threshold=0.5, resampler="multinomial",
                 path="auto", seed=0, device=None, shard_rank=0, shard_world=1, trial_offset=0):
        self.lib = _lib.load()
        self.hm = host_model
        self.N, self.B = int(num_particles), int(num_trials)
        self.dx, self.dq, self.dy = host_model.dx, host_model.dq, host_model.dy
        self.resampler = resampler
        self.device = _default_device() if device is None else int(device)
        cfg = _lib.PfConfig(self.N, self.B, float(ess_threshold), _RESAMPLERS[resampler], _PATHS[path],
                            int(seed) & 0xFFFFFFFFFFFFFFFF, self.device, int(trial_offset), int(shard_rank), int(shard_world))
        self.h = C.c_void_p()
        _lib.check(self.lib.bf_pf_create(C.byref(host_model.desc), C.byref(cfg), C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.bf_pf_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def launches(self):
        n = C.c_int64()
        _lib.check(self.lib.bf_pf_kernel_launches(self.h, C.byref(n)))
        return n.value

    def phase_times(self, reset=True):
        """Streamed path: ms spent in (propagate+search, stats+weights, tile-sum scan, cdf, final gather)."""
        out = (C.c_double * 5)()
        _lib.check(self.lib.bf_pf_phase_times(self.h, out, int(bool(reset))))
        return dict(zip(("propagate_search", "stats_weights", "scan_tiles", "cdf", "final_gather"), list(out)))

    # ------------------------------------------------------------------ shapes of the inputs
    def _in_shapes(self, T, per_trial):
        B, N = self.B, self.N
        return dict(emissions=((B, T, self.dy) if per_trial else (T, self.dy)), inputs=(T,),
                    init_particles=(B, N, self.dx), init_weights=(B, N), noise=(T, B, N, self.dq),
                    resample_u=((T, B) if self.resampler == "systematic" else (T, B, N)))

    def _io(self, T, per_trial, t_offset, resume, ins, outs):
        io = _lib.PfIO()
        io.T, io.t_offset, io.resume, io.emissions_per_trial = int(T), int(t_offset), int(bool(resume)), int(bool(per_trial))
        for k, v in ins.items():
            setattr(io, k, _ptr(v))
        for k, v in outs.items():
            setattr(io, k, _ptr(v))
        return io

    def run_host(self, emissions, inputs=None, init_particles=None, init_weights=None, noise=None, resample_u=None,
                 outputs=("weights", "particles"), t_offset=0, resume=False):
        """NumPy in / NumPy out.  Copies happen inside bf_pf_run_host; returns after the device has finished."""
        em = np.asarray(emissions, dtype=np.float32)
        per_trial = em.ndim == 3
        T = em.shape[-2] if em.ndim >= 2 else em.shape[0]
        shp = self._in_shapes(T, per_trial)
        ins = dict(emissions=_np32(em, shp["emissions"], "emissions"), inputs=_np32(inputs, shp["inputs"], "inputs"),
                   init_particles=_np32(init_particles, shp["init_particles"], "init_particles"),
                   init_weights=_np32(init_weights, shp["init_weights"], "init_weights"),
                   noise=_np32(noise, shp["noise"], "noise"), resample_u=_np32(resample_u, shp["resample_u"], "resample_u"))
        res, outs = {}, {}
        for name in outputs:
            if name in PF_OUTPUTS:
                field, shape, dt = PF_OUTPUTS[name]
                res[name] = _host_empty((self.B, T) + shape(self.N, self.dx), dt)
            elif name in PF_FINALS:
                field, shape = PF_FINALS[name]
                res[name] = _host_empty((self.B,) + shape(self.N, self.dx), np.float32)
            else:
                raise ValueError(f"unknown particle-filter output {name!r}")
            outs[field] = res[name]
        io = self._io(T, per_trial, t_offset, resume, ins, outs)
        _lib.check(self.lib.bf_pf_run_host(self.h, C.byref(io)))
        return res

    def run(self, emissions, inputs=None, init_particles=None, init_weights=None, noise=None, resample_u=None,
            outputs=("weights", "particles"), t_offset=0, resume=False, out_buffers=None):
        """torch CUDA tensors in / out, enqueued on torch's current stream, no synchronisation."""
        import torch
        dev = torch.device("cuda", self.device)
        f32 = lambda a, shape, name: None if a is None else self._t32(a, shape, name, dev)
        em = emissions
        per_trial = em.dim() == 3
        T = em.shape[-2]
        shp = self._in_shapes(T, per_trial)
        ins = dict(emissions=f32(em, shp["emissions"], "emissions"), inputs=f32(inputs, shp["inputs"], "inputs"),
                   init_particles=f32(init_particles, shp["init_particles"], "init_particles"),
                   init_weights=f32(init_weights, shp["init_weights"], "init_weights"),
                   noise=f32(noise, shp["noise"], "noise"), resample_u=f32(resample_u, shp["resample_u"], "resample_u"))
        res, outs = {}, {}
        for name in outputs:
            if name in PF_OUTPUTS:
                field, shape, dt = PF_OUTPUTS[name]
                full = (self.B, T) + shape(self.N, self.dx)
                tdt = torch.int32 if dt is np.int32 else torch.float32
            elif name in PF_FINALS:
                field, shape = PF_FINALS[name]
                full = (self.B,) + shape(self.N, self.dx)
                tdt = torch.float32
            else:
                raise ValueError(f"unknown particle-filter output {name!r}")
            if out_buffers is not None and name in out_buffers:
                buf = out_buffers[name]
                if tuple(buf.shape) != full or buf.dtype != tdt or not buf.is_contiguous():
                    raise ValueError(f"out_buffers[{name!r}] must be a contiguous {tdt} tensor of shape {full}")
            else:
                buf = torch.empty(full, dtype=tdt, device=dev)
            res[name] = buf
            outs[field] = buf
        io = self._io(T, per_trial, t_offset, resume, ins, outs)
        self._keep = (ins, outs)
        _lib.check(self.lib.bf_pf_run(self.h, C.byref(io), torch.cuda.current_stream(dev).cuda_stream))
        return res

    @staticmethod
    def _t32(a, shape, name, dev):
        import torch
        if not isinstance(a, torch.Tensor):
            a = torch.as_tensor(np.asarray(a, dtype=np.float32))
        a = a.to(device=dev, dtype=torch.float32).contiguous()
        if a.numel() != int(np.prod(shape)):
            raise ValueError(f"{name} has {a.numel()} elements, expected shape {tuple(shape)}")
        return a.view(shape)


class ShardedParticleFilterEngine(ParticleFilterEngine):
    """ONE bootstrap particle filter whose particles are sharded over the GPUs of a box: this rank owns
    `num_particles // world` consecutive particles.  Needs an initialised `torch.distributed` process group (one
    process per GPU, e.g. torchrun); the group carries only the set-up (CUDA-IPC handles, NCCL id).  Per time step
    libbfilt itself exchanges the weight statistics and the scan offsets with NCCL and gathers ancestors that live on
    another rank straight from its memory over NVLink (include/bfilt.h, bf_pf_shard_*).

    `run` / `run_host` take and return the LOCAL shard of every per-particle array (noise [T,1,N/world,dq],
    uniforms [T,1,N/world], outputs [1,T,N/world,...], ancestors as GLOBAL indices); per-step scalars (means,
    log_likelihood, ess, resampled) are those of the whole filter, written by rank 0 (other ranks get zeros)."""

    def __init__(self, host_model, num_particles, ess_threshold=0.5, resampler="multinomial", seed=0, device=None, group=None):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("ShardedParticleFilterEngine needs an initialised torch.distributed process group")
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if num_particles % world:
            raise ValueError(f"num_particles={num_particles} is not divisible by the {world} ranks")
        self.world, self.rank, self.N_global = world, rank, int(num_particles)
        super().__init__(host_model, num_particles // world, 1, ess_threshold, resampler, "streamed", seed, device,
                         shard_rank=rank, shard_world=world)
        if world == 1:
            return
        blob = C.create_string_buffer(_lib.SHARD_BLOB_BYTES)
        _lib.check(self.lib.bf_pf_shard_export(self.h, blob))
        blobs = [None] * world
        dist.all_gather_object(blobs, blob.raw, group=group)
        _lib.check(self.lib.bf_pf_shard_import(self.h, b"".join(blobs)))
        if self.lib.bf_pf_shard_comm_ready(self.h):       # the process already has a communicator for this (device, world, rank)
            return
        uid = [None]
        if rank == 0:
            buf = C.create_string_buffer(_lib.NCCL_ID_BYTES)
            _lib.check(self.lib.bf_nccl_unique_id(buf))
            uid = [buf.raw]
        dist.broadcast_object_list(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        _lib.check(self.lib.bf_pf_shard_comm_init(self.h, uid[0]))


class GaussianSumFilterEngine:
    """Owns a bf_gsf_handle: M components x B independent trials of one registry model."""

    def __init__(self, host_model, num_components, num_trials=1, device=None, flags=0, unscented=None, legacy=None):
        """unscented = (alpha, beta, kappa): the unscented Gaussian-sum filter (inference.py:379-456) instead of the EKF one.
        legacy = ("ukf", alpha, beta, kappa) or ("mcf", num_particles, seed): the moment approximations of the legacy
        gaussfilt.UKF / gaussfilt.MCF classes (gaussfilt.py:146-198)."""
        self.lib = _lib.load()
        self.hm = host_model
        self.M, self.B = int(num_components), int(num_trials)
        self.dx, self.dy = host_model.dx, host_model.dy
        self.device = _default_device() if device is None else int(device)
        cfg = _lib.GsfConfig(self.M, self.B, self.device, int(flags))
        self.h = C.c_void_p()
        _lib.check(self.lib.bf_gsf_create(C.byref(host_model.desc), C.byref(cfg), C.byref(self.h)))
        self.mc_np = 0
        if unscented is not None:
            a, b, k = (float(v) for v in unscented)
            _lib.check(self.lib.bf_gsf_set_unscented(self.h, a, b, k))
        if legacy is not None:
            if legacy[0] == "ukf":
                _lib.check(self.lib.bf_gsf_set_legacy_moments(self.h, _lib.LEGACY_UKF, float(legacy[1]), float(legacy[2]), float(legacy[3]), 0, 0))
            elif legacy[0] == "mcf":
                self.mc_np = int(legacy[1])
                _lib.check(self.lib.bf_gsf_set_legacy_moments(self.h, _lib.LEGACY_MCF, 0.0, 0.0, 0.0, self.mc_np, int(legacy[2]) & 0xFFFFFFFFFFFFFFFF))
            else:
                raise ValueError(f"unknown legacy moment approximation {legacy[0]!r}")

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.bf_gsf_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def launches(self):
        n = C.c_int64()
        _lib.check(self.lib.bf_gsf_kernel_launches(self.h, C.byref(n)))
        return n.value

    def _in_shapes(self, T, per_trial):
        B, M, dx = self.B, self.M, self.dx
        dq, dr = self.hm.desc.dq, self.hm.desc.dr
        return dict(emissions=((B, T, self.dy) if per_trial else (T, self.dy)), inputs=(T,), init_means=(B, M, dx),
                    init_covs=(B, M, dx, dx), init_weights=(B, M), Q_t=(T, dq, dq), R_t=(T, dr, dr),
                    mc_draws=(B, T, M, 2, max(self.mc_np, 1), dx))

    def _io(self, T, per_trial, resume, ins, outs):
        io = _lib.GsfIO()
        io.T, io.resume, io.emissions_per_trial = int(T), int(bool(resume)), int(bool(per_trial))
        for k, v in ins.items():
            setattr(io, k, _ptr(v))
        for k, v in outs.items():
            setattr(io, k, _ptr(v))
        return io

    def run_host(self, emissions, inputs=None, init_means=None, init_covs=None, init_weights=None,
                 outputs=("weights", "means", "covariances", "predicted_means", "predicted_covariances"), resume=False,
                 Q_t=None, R_t=None, mc_draws=None):
        """Q_t [T,dq,dq] / R_t [T,dr,dr]: time-varying noise covariances (`_get_params`, inference.py:21).
        mc_draws [B,T,M,2,np,dx]: the legacy Monte-Carlo filter's draws as offsets x - m (else Philox in the kernel)."""
        em = np.asarray(emissions, dtype=np.float32)
        per_trial = em.ndim == 3
        T = em.shape[-2]
        shp = self._in_shapes(T, per_trial)
        ins = {k: _np32(v, shp[k], k) for k, v in dict(emissions=em, inputs=inputs, init_means=init_means,
                                                        init_covs=init_covs, init_weights=init_weights,
                                                        Q_t=Q_t, R_t=R_t, mc_draws=mc_draws).items()}
        res, outs = {}, {}
        for name in outputs:
            if name not in GSF_OUTPUTS:
                raise ValueError(f"unknown Gaussian-sum-filter output {name!r}")
            field, shape = GSF_OUTPUTS[name]
            res[name] = _host_empty((self.B, T) + shape(self.M, self.dx), np.float32)
            outs[field] = res[name]
        io = self._io(T, per_trial, resume, ins, outs)
        _lib.check(self.lib.bf_gsf_run_host(self.h, C.byref(io)))
        return res

    def run(self, emissions, inputs=None, init_means=None, init_covs=None, init_weights=None,
            outputs=("weights", "means", "covariances", "predicted_means", "predicted_covariances"), resume=False,
            out_buffers=None, Q_t=None, R_t=None):
        import torch
        dev = torch.device("cuda", self.device)
        per_trial = emissions.dim() == 3
        T = emissions.shape[-2]
        shp = self._in_shapes(T, per_trial)
        t32 = ParticleFilterEngine._t32
        ins = {k: (None if v is None else t32(v, shp[k], k, dev))
               for k, v in dict(emissions=emissions, inputs=inputs, init_means=init_means, init_covs=init_covs,
                                init_weights=init_weights, Q_t=Q_t, R_t=R_t).items()}
        res, outs = {}, {}
        for name in outputs:
            if name not in GSF_OUTPUTS:
                raise ValueError(f"unknown Gaussian-sum-filter output {name!r}")
            field, shape = GSF_OUTPUTS[name]
            full = (self.B, T) + shape(self.M, self.dx)
            if out_buffers is not None and name in out_buffers:
                buf = out_buffers[name]
                if tuple(buf.shape) != full or buf.dtype != torch.float32 or not buf.is_contiguous():
                    raise ValueError(f"out_buffers[{name!r}] must be a contiguous float32 tensor of shape {full}")
            else:
                buf = torch.empty(full, dtype=torch.float32, device=dev)
            res[name] = buf
            outs[field] = buf
        io = self._io(T, per_trial, resume, ins, outs)
        self._keep = (ins, outs)
        _lib.check(self.lib.bf_gsf_run(self.h, C.byref(io), torch.cuda.current_stream(dev).cuda_stream))
        return res


AGSF_OUTPUTS = {
    # name: (io field, per-(trial,step) shape as a function of (M, L, dx), dtype)
    "weights": ("out_weights", lambda M, L, dx: (M,), np.float32),
    "means": ("out_means", lambda M, L, dx: (M, dx), np.float32),
    "covariances": ("out_covs", lambda M, L, dx: (M, dx, dx), np.float32),
    "log_likelihood": ("out_loglik", lambda M, L, dx: (), np.float32),
    "pre_weights": ("out_pre_weights", lambda M, L, dx: (L,), np.float32),
    "updated_means": ("out_updated_means", lambda M, L, dx: (L, dx), np.float32),
    "ancestors": ("out_ancestors", lambda M, L, dx: (M,), np.int32),
}


class AugmentedGaussianSumFilterEngine:
    """Owns a bf_agsf_handle: M components, each branched N1 x N2 ways per step, x B independent trials."""

    def __init__(self, host_model, num_components, num_trials=1, opt_args=(0.1, 0.1), seed=0, device=None):
        self.lib = _lib.load()
        self.hm = host_model
        self.M, self.N1, self.N2 = (int(v) for v in num_components)
        self.L = self.M * self.N1 * self.N2
        self.B = int(num_trials)
        self.dx, self.dy = host_model.dx, host_model.dy
        self.device = _default_device() if device is None else int(device)
        cfg = _lib.AgsfConfig(self.M, self.N1, self.N2, self.B, self.device, 0, float(opt_args[0]), float(opt_args[1]),
                              int(seed) & 0xFFFFFFFFFFFFFFFF)
        self.h = C.c_void_p()
        _lib.check(self.lib.bf_agsf_create(C.byref(host_model.desc), C.byref(cfg), C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.bf_agsf_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def run_host(self, emissions, inputs=None, init_means=None, init_covs=None, z_iid=None, s_iid=None, res_u=None,
                 outputs=("weights", "means", "covariances")):
        """z_iid [M,dx,N1] / s_iid [M*N1,dx,N2] / res_u [M]: one set of draws reused at every step (the reference), or with a
        leading T axis one per step (all three then)."""
        em = np.asarray(emissions, dtype=np.float32)
        per_trial = em.ndim == 3
        T = em.shape[-2]
        M, N1, N2, dx, B = self.M, self.N1, self.N2, self.dx, self.B
        per_step = any(a is not None and np.ndim(a) == nd + 1 for a, nd in ((z_iid, 3), (s_iid, 3), (res_u, 1)))
        TR = (T,) if per_step else ()
        ins = dict(emissions=_np32(em, (B, T, self.dy) if per_trial else (T, self.dy), "emissions"), inputs=_np32(inputs, (T,), "inputs"),
                   init_means=_np32(init_means, (B, M, dx), "init_means"), init_covs=_np32(init_covs, (B, M, dx, dx), "init_covs"),
                   z_iid=_np32(z_iid, TR + (M, dx, N1), "z_iid"), s_iid=_np32(s_iid, TR + (M * N1, dx, N2), "s_iid"),
                   res_u=_np32(res_u, TR + (M,), "res_u"))
        if ins["init_means"] is None:
            raise ValueError("init_means is required")
        io = _lib.AgsfIO()
        io.T, io.emissions_per_trial, io.rand_per_step = int(T), int(per_trial), int(per_step)
        for k, v in ins.items():
            setattr(io, k, _ptr(v))
        res = {}
        for name in outputs:
            if name not in AGSF_OUTPUTS:
                raise ValueError(f"unknown augmented-filter output {name!r}")
            field, shape, dt = AGSF_OUTPUTS[name]
            res[name] = _host_empty((B, T) + shape(M, self.L, dx), dt)
            setattr(io, field, _ptr(res[name]))
        _lib.check(self.lib.bf_agsf_run_host(self.h, C.byref(io)))
        return res


def optimal_resampling_device(weights, N, uniforms):
    """utils.optimal_resampling (utils.py:216-244) on the GPU: torch CUDA tensors weights [M], uniforms [M] -> (idx [N], w [N])."""
    import torch
    lib = _lib.load()
    w = weights.contiguous().float()
    u = uniforms.contiguous().float()
    idx = torch.empty(int(N), dtype=torch.int32, device=w.device)
    wo = torch.empty(int(N), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):         # the library launches on the current device
        _lib.check(lib.bf_optimal_resampling(w.data_ptr(), w.numel(), int(N), u.data_ptr(), idx.data_ptr(), wo.data_ptr(),
                                             torch.cuda.current_stream(w.device).cuda_stream))
    return idx, wo


def cumsum_canonical(w):
    """cumsum in lax.associative_scan order on the GPU (torch CUDA tensor in/out); bit-exact vs the oracle."""
    import torch
    lib = _lib.load()
    w = w.contiguous().float()
    out = torch.empty_like(w)
    with torch.cuda.device(w.device):
        _lib.check(lib.bf_cumsum(w.data_ptr(), out.data_ptr(), w.numel(), torch.cuda.current_stream(w.device).cuda_stream))
    return out


def resample_ancestors(w, u, resampler="multinomial"):
    """Ancestor indices of jax.random.choice(p=w) for supplied uniforms (torch CUDA tensors)."""
    import torch
    lib = _lib.load()
    w = w.contiguous().float()
    u = u.contiguous().float().reshape(-1)
    if resampler == "multinomial" and u.numel() != w.numel():
        raise ValueError("multinomial resampling needs one uniform per particle")
    anc = torch.empty(w.numel(), dtype=torch.int32, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(lib.bf_resample(w.data_ptr(), u.data_ptr(), anc.data_ptr(), w.numel(), _RESAMPLERS[resampler],
                                   torch.cuda.current_stream(w.device).cuda_stream))
    return anc
