"""ctypes binding of the C ABI in ``include/fabolas_b200.h``.

PyTorch tensors are only the buffer carrier: every call passes ``tensor.data_ptr()`` of FP64 CUDA
tensors to the hand-written sm_100a kernels in ``libfabolas_b200.so``.  There is NO CPU fallback:
importing the library fails loudly if it has not been built, and creating an :class:`Engine` fails
loudly without a CUDA device.

(Tests may hand an explicitly built emulator library to ``Engine(lib_path=..., device="cpu")`` to
replay the kernel logic without a GPU -- see ``tests/cusim``; nothing in this package does.)
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# FABOLAS_B200_LIB: development override (tools/ build kernel variants next to the shipped library and time them)
DEFAULT_LIB = os.environ.get("FABOLAS_B200_LIB") or os.path.join(_HERE, "_lib", "libfabolas_b200.so")

FAB_FAMILY_M52 = 0
FAB_FAMILY_M52_BASIS = 1
FAB_PRIOR_FABOLAS = 0      # fabolas.log_likelihood's prior (fabolas.py:27-50)
FAB_PRIOR_ES = 1           # entropy_search's log_prob prior (entropy_search.py:49-69)

_c_int = ctypes.c_int
_c_dbl = ctypes.c_double
_vp = ctypes.c_void_p

# name -> (restype, argtypes); also the list of symbols include/fabolas_b200.h declares.
SIGNATURES = {
    "fab_version": (_c_int, []),
    "fab_ctx_create": (_c_int, [ctypes.POINTER(_vp), _c_int, _vp]),
    "fab_ctx_destroy": (_c_int, [_vp]),
    "fab_last_error": (ctypes.c_char_p, [_vp]),
    "fab_ctx_set_timing": (_c_int, [_vp, _c_int]),
    "fab_ctx_get_timing": (_c_int, [_vp, ctypes.POINTER(_c_dbl), _c_int]),
    "fab_gp_set_data": (_c_int, [_vp, _c_int, _vp, _c_int, _c_int, _vp, _c_dbl, _c_dbl]),
    "fab_gp_loglik_batch": (_c_int, [_vp, _vp, _vp, _c_int, _vp, _vp, _vp]),
    "fab_gp_loglik_batch_host": (_c_int, [_vp, _vp, _vp, _c_int, _vp, _vp, _vp]),
    "fab_ctx_synchronize": (_c_int, [_vp]),
    "fab_gp_factorize_batch": (_c_int, [_vp, _vp, _vp, _c_int, _vp, _vp]),
    "fab_gp_fit_batch": (_c_int, [_vp, _vp, _vp, _c_int, _vp, _vp]),
    "fab_gp_predict_batch": (_c_int, [_vp, _vp, _c_int, _c_int, _vp, _vp, _vp]),
    "fab_epmgp_joint_min_batch": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fab_es_setup": (_c_int, [_vp, _vp, _c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int]),
    "fab_es_information_gain_batch": (_c_int, [_vp, _vp, _c_int, _vp, _vp, _vp, _vp]),
    "fab_kernel_matrix": (_c_int, [_vp, _c_int, _c_int, _c_dbl, _vp, _c_int, _vp, _c_int, _vp, _c_int, _c_int, _c_dbl, _vp, _vp]),
    "fab_gp_prior_draw_batch": (_c_int, [_vp, _vp, _c_int, ctypes.c_ulonglong, _vp, _vp, _vp]),
    "fab_ei_batch": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _vp, _c_dbl, _vp, _vp, _vp]),
    "fab_peer_alloc": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp]),
    "fab_peer_connect": (_c_int, [_vp, _vp]),
    "fab_gp_loglik_exchange": (_c_int, [_vp, _vp, _vp, _c_int, _vp, _vp, _vp]),
    "fab_peer_wait": (_c_int, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_c_int), ctypes.POINTER(_c_int)]),
    "fab_peer_error": (_c_int, [_vp]),
    "fab_peer_free": (_c_int, [_vp]),
    "fab_mcmc_run": (_c_int, [_vp, _c_int, _vp, _vp, _c_int, _c_int, _c_int, _c_int, ctypes.c_ulonglong,
                              ctypes.c_longlong, _c_dbl, _vp, _vp, _vp]),
    "fab_mcmc_peer_alloc": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _vp]),
    "fab_mcmc_peer_connect": (_c_int, [_vp, _vp]),
    "fab_mcmc_peer_state": (_c_int, [_vp, _vp, _vp, _c_int]),
    "fab_mcmc_run_sharded": (_c_int, [_vp, _c_int, _c_int, _c_int, ctypes.c_ulonglong, ctypes.c_longlong, _c_dbl, _vp]),
    "fab_mcmc_peer_error": (_c_int, [_vp]),
    "fab_mcmc_peer_free": (_c_int, [_vp]),
    "fab_mcmc_log_prior": (_c_int, [_vp, _c_int, _vp, _c_int, _c_int, _vp, _vp, _vp]),
    "fab_ctx_debug_set_mcmc_ctas": (_c_int, [_vp, _c_int]),
    "fab_ctx_debug_set_max_slots": (_c_int, [_vp, _c_int]),
    "fab_ctx_debug_force_slices": (_c_int, [_vp, _c_int]),
    "fab_ctx_debug_set_predict_chunk": (_c_int, [_vp, _c_int]),
    "fab_gp_debug_get_factor": (_c_int, [_vp, _c_int, _vp]),
    "fab_debug_dag_verify": (_c_int, [_c_int, _c_int, _c_int, ctypes.c_char_p, _c_int, ctypes.POINTER(_c_int)]),
}


class FabolasError(RuntimeError):
    pass


_LIBS: dict = {}


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise FabolasError(
            f"{path} is missing: build the CUDA extension first (python -m fabolas_b200.build). "
            "fabolas_b200 has no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _LIBS[path] = lib
    return lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else _vp(t.data_ptr())


class Engine:
    """One ``fab_ctx``: a device, a stream and the resident GP problem / factorised models."""

    def __init__(self, device: Optional[str] = None, lib_path: Optional[str] = None):
        self.lib = load_library(lib_path)
        if lib_path is None:
            if not torch.cuda.is_available():
                raise FabolasError("fabolas_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
            self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
            if self.device.type != "cuda":
                raise FabolasError("fabolas_b200 engines run on CUDA devices only")
            index = self.device.index if self.device.index is not None else torch.cuda.current_device()
            self.device = torch.device("cuda", index)
            with torch.cuda.device(index):
                stream = torch.cuda.current_stream().cuda_stream
        else:  # emulator library handed in by a test
            self.device = torch.device(device or "cpu")
            index, stream = 0, 0
        self._ctx = _vp()
        rc = self.lib.fab_ctx_create(ctypes.byref(self._ctx), index, _vp(stream))
        if rc != 0:
            raise FabolasError(f"fab_ctx_create failed with {rc}")
        self.family = None
        self.N = self.D = self.d = self.Pk = 0

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.fab_ctx_destroy(self._ctx)
            self._ctx = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.fab_last_error(self._ctx)
            raise FabolasError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def tensor(self, a, dtype=torch.float64) -> torch.Tensor:
        """Move/convert `a` to a contiguous tensor of `dtype` on the engine's device."""
        t = torch.as_tensor(a, dtype=dtype)
        return t.to(self.device).contiguous()

    def empty(self, *shape, dtype=torch.float64) -> torch.Tensor:
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def set_timing(self, enable: bool):
        self._check(self.lib.fab_ctx_set_timing(self._ctx, int(enable)), "fab_ctx_set_timing")

    def debug_set_max_slots(self, n: int):
        self._check(self.lib.fab_ctx_debug_set_max_slots(self._ctx, n), "fab_ctx_debug_set_max_slots")

    def debug_set_predict_chunk(self, jc: int):
        self._check(self.lib.fab_ctx_debug_set_predict_chunk(self._ctx, int(jc)), "fab_ctx_debug_set_predict_chunk")

    def debug_force_slices(self, on: bool):
        self._check(self.lib.fab_ctx_debug_force_slices(self._ctx, int(on)), "fab_ctx_debug_force_slices")

    def get_timing(self):
        """Durations (ms) of the kernels launched by the most recent batched call."""
        buf = (_c_dbl * 8)()
        n = self.lib.fab_ctx_get_timing(self._ctx, buf, 8)
        if n < 0:
            self._check(n, "fab_ctx_get_timing")
        return [buf[i] for i in range(n)]

    # ------------------------------------------------------------------------------------------
    def set_data(self, family: int, X, y, mean: float, amp_mult: Optional[float] = None):
        X = self.tensor(X)
        y = self.tensor(y).reshape(-1)
        if X.dim() != 2 or y.numel() != X.shape[0]:
            raise ValueError("X must be (N, D) and y (N,)")
        N, D = X.shape
        amp_mult = float(D if amp_mult is None else amp_mult)
        self._check(self.lib.fab_gp_set_data(self._ctx, family, _ptr(X), N, D, _ptr(y), float(mean), amp_mult),
                    "fab_gp_set_data")
        self.family, self.N, self.D = family, N, D
        self.d = D - 1 if family == FAB_FAMILY_M52_BASIS else D
        self.Pk = self.d + 3 if family == FAB_FAMILY_M52_BASIS else self.d + 1

    def _theta(self, theta, yerr2):
        theta = self.tensor(theta)
        if theta.dim() == 1:
            theta = theta[None]
        if theta.shape[1] != self.Pk:
            raise ValueError(f"theta must have {self.Pk} columns, got {theta.shape[1]}")
        yerr2 = self.tensor(yerr2).reshape(-1)
        if yerr2.numel() == 1 and theta.shape[0] > 1:
            yerr2 = yerr2.expand(theta.shape[0]).contiguous()
        if yerr2.numel() != theta.shape[0]:
            raise ValueError("one yerr2 per theta row")
        return theta, yerr2

    def loglik(self, theta, yerr2, grad: bool = False, out=None):
        """ll (B,), [grad (B, Pk + 1),] info (B,) of B hyper-parameter rows: one launch of the fused kernel.
        `out` = (ll, grad or None, info): caller-owned contiguous result tensors (e.g. views of one packed
        buffer that a single all-gather ships)."""
        theta, yerr2 = self._theta(theta, yerr2)
        B = theta.shape[0]
        if out is not None:
            ll, g, info = out
            if ll.numel() != B or info.numel() != B or (grad and (g is None or g.numel() != B * (self.Pk + 1))):
                raise ValueError("out tensors do not match the batch")
            if not (ll.is_contiguous() and info.is_contiguous() and (g is None or g.is_contiguous())):
                raise ValueError("out tensors must be contiguous")
        else:
            ll = self.empty(B)
            info = self.empty(B, dtype=torch.int32)
            g = self.empty(B, self.Pk + 1) if grad else None
        self._check(self.lib.fab_gp_loglik_batch(self._ctx, _ptr(theta), _ptr(yerr2), B, _ptr(ll), _ptr(g),
                                                 _ptr(info)), "fab_gp_loglik_batch")
        return (ll, g, info) if grad else (ll, info)

    def loglik_host(self, theta, yerr2, grad: bool = False, out=None):
        """loglik() for HOST tensors (float64, contiguous; info int32): pinned tensors are read and written by the kernel
        itself over the bus, pageable ones are staged by the library.  Enqueued on the engine's stream: synchronise
        (Engine.synchronize or torch.cuda.synchronize) before reading the outputs.  Returns (ll, [grad,] info)."""
        if theta.device.type != "cpu" or yerr2.device.type != "cpu":
            raise ValueError("loglik_host takes host tensors; use loglik for device tensors")
        if theta.dtype != torch.float64 or yerr2.dtype != torch.float64 or not theta.is_contiguous() or not yerr2.is_contiguous():
            raise ValueError("float64 contiguous host tensors expected")
        if theta.dim() != 2 or theta.shape[1] != self.Pk or yerr2.numel() != theta.shape[0]:
            raise ValueError(f"theta must be B x {self.Pk}, yerr2 B")
        B = theta.shape[0]
        if out is not None:
            ll, g, info = out
        else:
            pin = theta.is_pinned()
            ll = torch.empty(B, dtype=torch.float64, pin_memory=pin)
            info = torch.empty(B, dtype=torch.int32, pin_memory=pin)
            g = torch.empty(B, self.Pk + 1, dtype=torch.float64, pin_memory=pin) if grad else None
        if ll.numel() != B or info.numel() != B or info.dtype != torch.int32 or (grad and (g is None or g.numel() != B * (self.Pk + 1))):
            raise ValueError("out tensors do not match the batch")
        self._check(self.lib.fab_gp_loglik_batch_host(self._ctx, _ptr(theta), _ptr(yerr2), B, _ptr(ll), _ptr(g) if grad else None,
                                                      _ptr(info)), "fab_gp_loglik_batch_host")
        return (ll, g, info) if grad else (ll, info)

    def synchronize(self):
        self._check(self.lib.fab_ctx_synchronize(self._ctx), "fab_ctx_synchronize")

    def factorize(self, theta, yerr2):
        theta, yerr2 = self._theta(theta, yerr2)
        B = theta.shape[0]
        ll = self.empty(B)
        info = self.empty(B, dtype=torch.int32)
        self._check(self.lib.fab_gp_factorize_batch(self._ctx, _ptr(theta), _ptr(yerr2), B, _ptr(ll), _ptr(info)),
                    "fab_gp_factorize_batch")
        self.B = B
        return ll, info

    def fit(self, theta, yerr2):
        """Make B models resident (factor, inverse factor, alpha).  Returns (ll, info)."""
        theta, yerr2 = self._theta(theta, yerr2)
        B = theta.shape[0]
        ll = self.empty(B)
        info = self.empty(B, dtype=torch.int32)
        self._check(self.lib.fab_gp_fit_batch(self._ctx, _ptr(theta), _ptr(yerr2), B, _ptr(ll), _ptr(info)),
                    "fab_gp_fit_batch")
        self.B = B
        return ll, info

    def predict(self, T, want_var: bool = True, want_cov: bool = False):
        """mu (B, M) [, var (B, M)] [, cov (B, M, M)] of every resident model at the rows of T."""
        T = self.tensor(T)
        if T.dim() == 1:
            T = T[None]
        per_model = T.dim() == 3          # (B, M, D): every model predicts at its own points
        if T.shape[-1] != self.D or (per_model and T.shape[0] != self.B):
            raise ValueError(f"T must be (M, {self.D}) or ({self.B}, M, {self.D})")
        M = T.shape[-2]
        mu = self.empty(self.B, M)
        var = self.empty(self.B, M) if want_var else None
        cov = self.empty(self.B, M, M) if want_cov else None
        self._check(self.lib.fab_gp_predict_batch(self._ctx, _ptr(T), int(per_model), M, _ptr(mu), _ptr(var), _ptr(cov)),
                    "fab_gp_predict_batch")
        out = [mu]
        if want_var:
            out.append(var)
        if want_cov:
            out.append(cov)
        return tuple(out)

    def prior_draw_max(self, reps, seed: int, want_sample: bool = False):
        """max over the points of ONE prior draw GP.sample(reps[b]) per resident model (fabolas.py:198-199):
        reps (B, Z, D) -> y_max (B,) [, sample (B, Z)], info (B,)."""
        reps = self.tensor(reps)
        if reps.dim() != 3 or reps.shape[0] != self.B or reps.shape[2] != self.D:
            raise ValueError(f"reps must be ({self.B}, Z, {self.D})")
        Z = reps.shape[1]
        out = self.empty(self.B)
        sample = self.empty(self.B, Z) if want_sample else None
        info = self.empty(self.B, dtype=torch.int32)
        self._check(self.lib.fab_gp_prior_draw_batch(self._ctx, _ptr(reps), Z, int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(out),
                                                     _ptr(sample), _ptr(info)), "fab_gp_prior_draw_batch")
        return (out, sample, info) if want_sample else (out, info)

    def expected_improvement(self, mu, var, y_max, xi: float = 0.0, want_ei: bool = True):
        """EI (B, M), best value (B,), argmax (B,) for predictive moments mu/var (B, M)."""
        mu, var = self.tensor(mu), self.tensor(var)
        if mu.dim() == 1:
            mu, var = mu[None], var[None]
        B, M = mu.shape
        y_max = self.tensor(y_max).reshape(-1)
        if y_max.numel() == 1 and B > 1:
            y_max = y_max.expand(B).contiguous()
        ei = self.empty(B, M) if want_ei else None
        bv = self.empty(B)
        bi = self.empty(B, dtype=torch.int64)
        self._check(self.lib.fab_ei_batch(self._ctx, _ptr(mu), _ptr(var), B, M, _ptr(y_max), float(xi), _ptr(ei),
                                          _ptr(bv), _ptr(bi)), "fab_ei_batch")
        return ei, bv, bi

    def joint_min(self, mu, Sigma, with_derivatives: bool = True, return_sweeps: bool = False):
        """Batched epmgp.joint_min: mu (B, Z), Sigma (B, Z, Z) -> logP [, dMu, dSigma, dMuMu], info."""
        mu, Sigma = self.tensor(mu), self.tensor(Sigma)
        if mu.dim() == 1:
            mu, Sigma = mu[None], Sigma[None]
        B, Z = mu.shape
        T = Z * (Z + 1) // 2
        logP = self.empty(B, Z)
        info = self.empty(B, dtype=torch.int32)
        sweeps = self.empty(B, Z, dtype=torch.int32) if return_sweeps else None
        dMu = self.empty(B, Z, Z) if with_derivatives else None
        dSg = self.empty(B, Z, T) if with_derivatives else None
        dMM = self.empty(B, Z, Z, Z) if with_derivatives else None
        self._check(self.lib.fab_epmgp_joint_min_batch(self._ctx, _ptr(mu), _ptr(Sigma), B, Z, _ptr(logP), _ptr(dMu),
                                                       _ptr(dSg), _ptr(dMM), _ptr(info), _ptr(sweeps)),
                    "fab_epmgp_joint_min_batch")
        out = (logP, dMu, dSg, dMM, info) if with_derivatives else (logP, info)
        return out + (sweeps,) if return_sweeps else out

    def es_setup(self, reps, cov_reps, p_min, U, Omega):
        """Entropy-Search state of the resident models.  reps (B, Z, D); cov_reps (B, Z, Z) unclipped;
        p_min = (logP, dMu, dSigma, dMuMu) batched; U (B, Z); Omega (B, I, Z)."""
        reps, cov_reps, U, Omega = (self.tensor(t) for t in (reps, cov_reps, U, Omega))
        logP, dMu, dSg, dMM = (self.tensor(t) for t in p_min[:4])
        B, Z, D = reps.shape
        if B != self.B or D != self.D or Omega.shape[0] != B or Omega.shape[2] != Z:
            raise ValueError("shape mismatch in es_setup")
        self._es_keep = (reps, cov_reps, U, Omega, logP, dMu, dSg, dMM)     # alive until the stream has consumed them
        self._check(self.lib.fab_es_setup(self._ctx, _ptr(reps), Z, _ptr(cov_reps), _ptr(logP), _ptr(dMu), _ptr(dSg),
                                          _ptr(dMM), _ptr(U), _ptr(Omega), Omega.shape[1]), "fab_es_setup")

    def information_gain(self, Xc, return_mixture: bool = False):
        """Information gain of every row of Xc (M, D) -> ig (M,), nan_flag (M,) [, mix_mean, mix_var]."""
        Xc = self.tensor(Xc)
        if Xc.dim() == 1:
            Xc = Xc[None]
        M = Xc.shape[0]
        ig = self.empty(M)
        flag = self.empty(M, dtype=torch.int32)
        mm = self.empty(M) if return_mixture else None
        mv = self.empty(M) if return_mixture else None
        self._check(self.lib.fab_es_information_gain_batch(self._ctx, _ptr(Xc), M, _ptr(ig), _ptr(flag), _ptr(mm), _ptr(mv)),
                    "fab_es_information_gain_batch")
        return (ig, flag, mm, mv) if return_mixture else (ig, flag)

    def kernel_matrix(self, family: int, theta, Xa, Xb=None, amp_mult: Optional[float] = None, nu_code: int = 0,
                      grad: bool = False, nu: float = 0.0):
        """K (B, Na, Nb) [, dK (B, Na, Nb, Pk)] of k(Xa, Xb) for every row of theta."""
        Xa = self.tensor(Xa)
        Xb = Xa if Xb is None else self.tensor(Xb)
        theta = self.tensor(theta)
        if theta.dim() == 1:
            theta = theta[None]
        B, Pk = theta.shape
        Na, D = Xa.shape
        Nb = Xb.shape[0]
        amp_mult = float(D if amp_mult is None else amp_mult)
        K = self.empty(B, Na, Nb)
        dK = self.empty(B, Na, Nb, Pk) if grad else None
        self._check(self.lib.fab_kernel_matrix(self._ctx, family, nu_code, float(nu), _ptr(theta), B, _ptr(Xa), Na, _ptr(Xb), Nb, D,
                                               amp_mult, _ptr(K), _ptr(dK)), "fab_kernel_matrix")
        return (K, dK) if grad else K

    # ------------------------------------------------------------------------------------------
    # multi-GPU likelihood step with the exchange fused into the kernel (peer memory over NVLink / NVSwitch)
    def peer_setup(self, rows_per_rank: int, group=None):
        """Allocate this rank's gather buffers, exchange the CUDA IPC handles over torch.distributed and map the
        peers.  Without an initialised process group: a world of one (the local rank is its own peer)."""
        import torch.distributed as dist
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        rank, world = (dist.get_rank(group), dist.get_world_size(group)) if multi else (0, 1)
        h = (ctypes.c_ubyte * 64)()
        rc = self.lib.fab_peer_alloc(self._ctx, rank, world, int(rows_per_rank), ctypes.cast(h, _vp))
        buf = self._exchange_handles(h, group)                # every rank takes part, whatever its own outcome
        if rc == 0:
            rc = self.lib.fab_peer_connect(self._ctx, None if buf is None else ctypes.cast(buf, _vp))
        self._agree(rc, "fab_peer_alloc / fab_peer_connect", group, self.lib.fab_peer_free)
        self._peer_views = {}
        self.peer_world, self.peer_rank, self.peer_rows = world, rank, int(rows_per_rank)

    def loglik_exchange(self, theta, yerr2, grad: bool = False, out=None):
        """loglik() on this rank's walkers whose result rows also land in every rank's gather buffer.  theta / yerr2 may be
        PINNED host tensors (B x Pk, B; float64, contiguous): the kernel then fetches them over the bus itself."""
        if isinstance(theta, torch.Tensor) and theta.device.type == "cpu" and theta.is_pinned() and self.device.type == "cuda":
            if not (isinstance(yerr2, torch.Tensor) and yerr2.is_pinned() and theta.dtype == yerr2.dtype == torch.float64
                    and theta.is_contiguous() and yerr2.is_contiguous() and theta.dim() == 2 and theta.shape[1] == self.Pk
                    and yerr2.numel() == theta.shape[0]):
                raise ValueError("pinned host inputs: float64 contiguous theta (B x Pk) and yerr2 (B), both pinned")
        else:
            theta, yerr2 = self._theta(theta, yerr2)
        B = theta.shape[0]
        if out is not None:
            ll, g, info = out if grad else (out[0], None, out[-1])
        else:
            ll = self.empty(B)
            info = self.empty(B, dtype=torch.int32)
            g = self.empty(B, self.Pk + 1) if grad else None
        self._check(self.lib.fab_gp_loglik_exchange(self._ctx, _ptr(theta), _ptr(yerr2), B, _ptr(ll), _ptr(g), _ptr(info)),
                    "fab_gp_loglik_exchange")
        return (ll, g, info) if grad else (ll, info)

    def peer_wait(self) -> torch.Tensor:
        """Enqueue the wait for every rank's rows of the last exchanged step; returns the (world * B, Pk + 3) view of
        this rank's gather buffer: [grad (Pk+1) | ll | info] per walker, rank-major."""
        ptr, rows, rd = _vp(), _c_int(), _c_int()
        self._check(self.lib.fab_peer_wait(self._ctx, ctypes.byref(ptr), ctypes.byref(rows), ctypes.byref(rd)), "fab_peer_wait")
        key = ptr.value
        if key not in self._peer_views:
            shape = (rows.value, rd.value)
            if self.device.type == "cuda":
                class _Raw:
                    __cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (key, False), "version": 3,
                                                "strides": None}
                self._peer_views[key] = torch.as_tensor(_Raw(), device=self.device)
            else:
                import numpy as np
                arr = np.ctypeslib.as_array(ctypes.cast(key, ctypes.POINTER(_c_dbl)), shape=shape)
                self._peer_views[key] = torch.from_numpy(arr)
        return self._peer_views[key]

    def peer_error(self) -> bool:
        return bool(self.lib.fab_peer_error(self._ctx))

    def peer_free(self):
        self._peer_views = {}
        self.peer_world = self.peer_rows = 0
        self._check(self.lib.fab_peer_free(self._ctx), "fab_peer_free")

    # ------------------------------------------------------------------------------------------
    def mcmc_run(self, prior: int, coords, nsteps: int, log_prob=None, seed: int = 0, iter0: int = 0, a: float = 2.0,
                 store_chain: bool = True, naccepted=None):
        """emcee's run_mcmc on the device (one persistent kernel): returns (coords, log_prob, chain, chain_lp,
        naccepted); coords (K, P) and log_prob are advanced IN PLACE when they are tensors on the device."""
        coords = self.tensor(coords)
        if coords.dim() != 2:
            raise FabolasError("coords must be (nwalkers, ndim)")
        K, P = coords.shape
        init = log_prob is None
        lp = self.empty(K) if init else self.tensor(log_prob)
        chain = self.empty(nsteps, K, P) if store_chain else None
        chain_lp = self.empty(nsteps, K) if store_chain else None
        nacc = torch.zeros(K, dtype=torch.int64, device=self.device) if naccepted is None else naccepted
        self._check(self.lib.fab_mcmc_run(self._ctx, int(prior), _ptr(coords), _ptr(lp), K, P, int(nsteps), int(init),
                                          int(seed) & 0xFFFFFFFFFFFFFFFF, int(iter0), float(a), _ptr(chain),
                                          _ptr(chain_lp), _ptr(nacc)), "fab_mcmc_run")
        return coords, lp, chain, chain_lp, nacc

    def _exchange_handles(self, h, group):
        """(rank, world, all handles as a ctypes buffer or None) for a 64-byte CUDA IPC handle of this rank."""
        import torch.distributed as dist
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if not multi:
            return None
        world = dist.get_world_size(group)
        mine = torch.tensor(list(h), dtype=torch.uint8, device=self.device)
        allh = torch.empty(world * 64, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(allh, mine, group=group)
        return (ctypes.c_ubyte * (64 * world))(*allh.cpu().tolist())

    def _agree(self, rc: int, what: str, group, undo):
        """Every rank learns whether every rank mapped its peers (a rank that failed must not leave the others in a
        collective): all succeed, or all release what they mapped and raise."""
        import torch.distributed as dist
        msg = self.lib.fab_last_error(self._ctx).decode() if rc != 0 else ""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            ok = torch.tensor([1.0 if rc == 0 else 0.0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if ok.item() == 0.0:
                undo(self._ctx)
                raise FabolasError(f"{what} failed on at least one rank" + (f" (this rank: {rc}: {msg})" if rc != 0 else ""))
        elif rc != 0:
            undo(self._ctx)
            raise FabolasError(f"{what} failed ({rc}): {msg}")

    def mcmc_peer_setup(self, K: int, P: int, group=None):
        """Replicas of a K-walker ensemble on every rank of `group` (default: the world), mapped into each other."""
        import torch.distributed as dist
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        rank, world = (dist.get_rank(group), dist.get_world_size(group)) if multi else (0, 1)
        h = (ctypes.c_ubyte * 64)()
        rc = self.lib.fab_mcmc_peer_alloc(self._ctx, rank, world, int(K), int(P), ctypes.cast(h, _vp))
        buf = self._exchange_handles(h, group)                # every rank takes part, whatever its own outcome
        if rc == 0:
            rc = self.lib.fab_mcmc_peer_connect(self._ctx, None if buf is None else ctypes.cast(buf, _vp))
        self._agree(rc, "fab_mcmc_peer_alloc / fab_mcmc_peer_connect", group, self.lib.fab_mcmc_peer_free)
        self.mcmc_world, self.mcmc_rank, self.mcmc_K, self.mcmc_P = world, rank, int(K), int(P)

    def mcmc_run_sharded(self, prior: int, coords, nsteps: int, log_prob=None, seed: int = 0, iter0: int = 0,
                         a: float = 2.0, naccepted=None):
        """mcmc_run with the walkers sharded over the ranks set up by mcmc_peer_setup; every rank passes the SAME
        coords (and log_prob) and gets the same (coords, log_prob) back; naccepted counts this rank's proposals."""
        coords = self.tensor(coords)
        K, P = coords.shape
        if (K, P) != (getattr(self, "mcmc_K", 0), getattr(self, "mcmc_P", 0)):
            raise FabolasError("mcmc_peer_setup was made for a different ensemble shape")
        init = log_prob is None
        lp = self.empty(K) if init else self.tensor(log_prob)
        nacc = torch.zeros(K, dtype=torch.int64, device=self.device) if naccepted is None else naccepted
        self._check(self.lib.fab_mcmc_peer_state(self._ctx, _ptr(coords), None if init else _ptr(lp), 1), "fab_mcmc_peer_state")
        self._check(self.lib.fab_mcmc_run_sharded(self._ctx, int(prior), int(nsteps), int(init), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                  int(iter0), float(a), _ptr(nacc)), "fab_mcmc_run_sharded")
        self._check(self.lib.fab_mcmc_peer_state(self._ctx, _ptr(coords), _ptr(lp), 0), "fab_mcmc_peer_state")
        return coords, lp, nacc

    def mcmc_peer_error(self) -> bool:
        return bool(self.lib.fab_mcmc_peer_error(self._ctx))

    def mcmc_peer_free(self):
        self.mcmc_world = self.mcmc_K = self.mcmc_P = 0
        self._check(self.lib.fab_mcmc_peer_free(self._ctx), "fab_mcmc_peer_free")

    def log_prior(self, prior: int, params, want_theta: bool = False):
        """Log-prior of the reference's sampler log-probabilities for rows of walker positions (K, P)."""
        p = self.tensor(params)
        K, P = p.shape
        out = self.empty(K)
        theta = torch.zeros(K, P - 1, dtype=torch.float64, device=self.device) if want_theta else None
        yerr2 = torch.zeros(K, dtype=torch.float64, device=self.device) if want_theta else None
        self._check(self.lib.fab_mcmc_log_prior(self._ctx, int(prior), _ptr(p), K, P, _ptr(out), _ptr(theta),
                                                _ptr(yerr2)), "fab_mcmc_log_prior")
        return (out, theta, yerr2) if want_theta else out

    def debug_set_mcmc_ctas(self, n: int):
        self._check(self.lib.fab_ctx_debug_set_mcmc_ctas(self._ctx, int(n)), "fab_ctx_debug_set_mcmc_ctas")

    def debug_factor(self, b: int) -> torch.Tensor:
        L = self.empty(self.N, self.N)
        self._check(self.lib.fab_gp_debug_get_factor(self._ctx, b, _ptr(L)), "fab_gp_debug_get_factor")
        return L
