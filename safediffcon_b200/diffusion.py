"""Diffusion process: drop-in for GaussianDiffusion (/root/reference/1D/model/diffusion.py:21-746) on the
sampling path -- ``sample`` / ``p_sample_loop`` / ``ddim_sample`` / ``p_sample`` / ``model_predictions`` keep the
reference signatures and keyword contract; each reverse step is ONE fused CUDA launch (csrc/posterior.cu) next
to the denoiser evaluation, with no host synchronisation inside the chain.

Scope: objective 'pred_noise', temporal 2-D layout [B, C, Nt, Nx], u0/uT/w conditioning,
pad writes, safety guidance.  Options that no shipped 1D config enables (self-conditioning, residual
conditioning, two-model evaluation, recurrence, expand_condition, pred_x0/pred_v) raise NotImplementedError.
``forward`` / ``p_losses`` (training loss) and the ``enable_grad`` last DDIM step run the denoiser's CUDA forward with its
full CUDA backward (parameter gradients), so the reference's InferenceFT / PostTrainPipeline optimiser loops work on it.
"""
import ctypes
import math
import os
import types
from collections import OrderedDict, namedtuple

import torch
import torch.nn as nn

from . import _lib as L
from .guidance import SafetyGuidance

ModelPrediction = namedtuple('ModelPrediction', ['pred_noise', 'pred_x_start'])

SAMPLER_DDIM, SAMPLER_DDPM = 0, 1
_INIT_TAG = 0x7FFFFFFF  # RNG tag of the initial x_T draw
# Chains replay ONE captured CUDA graph per reverse step (the C++ executor keeps every activation in a workspace, so a step is
# capturable at any batch size; below ~256 samples the ~125 launches of a step cost more host time than device time).
# SDC_GRAPH_MAX_BATCH=0 disables graphs, a positive value restricts them to batches up to that size.
GRAPH_MAX_BATCH = int(os.environ.get("SDC_GRAPH_MAX_BATCH", str(1 << 30)))
GRAPH_CACHE_ENTRIES = 4


class _GraphCache:
    """Captured reverse-step graphs of one GaussianDiffusion; never deep-copied (EMA wrappers copy the module)."""

    def __init__(self):
        self.entries = OrderedDict()

    def __deepcopy__(self, memo):
        return _GraphCache()


def linear_beta_schedule(timesteps):
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)


def cosine_beta_schedule(timesteps, s=0.008):
    """Nichol & Dhariwal cosine schedule in fp64 (reference model_utils.py:148-158)."""
    k = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
    abar = torch.cos(((k / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    abar = abar / abar[0]
    return torch.clip(1 - (abar[1:] / abar[:-1]), 0, 0.999)


def extract(a, t, x_shape):
    b, *_ = t.shape
    return a.gather(-1, t).reshape(b, *((1,) * (len(x_shape) - 1)))


class GaussianDiffusion(nn.Module):
    def __init__(self, model, *, seq_length, timesteps=1000, sampling_timesteps=None, objective='pred_noise',
                 beta_schedule='cosine', ddim_sampling_eta=0., auto_normalize=False, guidance_u0=True,
                 conditioned_on_residual=None, residual_on_u0=False, temporal=False, use_conv2d=False,
                 is_condition_u0=False, is_condition_uT=False, is_condition_u0_zero_pred_noise=True,
                 is_condition_uT_zero_pred_noise=True, condition_idx=10, recurrence=False, recurrence_k=1,
                 normalize_beta=False, train_on_padded_locations=False, train_on_partially_observed=None,
                 set_unobserved_to_zero_during_sampling=False, is_model_w=False, eval_two_models=False,
                 expand_condition=False, prior_beta=1):
        super().__init__()
        unsupported = dict(conditioned_on_residual=conditioned_on_residual, recurrence=recurrence, is_model_w=is_model_w,
                           eval_two_models=eval_two_models, expand_condition=expand_condition, auto_normalize=auto_normalize,
                           set_unobserved_to_zero_during_sampling=set_unobserved_to_zero_during_sampling)
        bad = [k for k, v in unsupported.items() if v]
        if bad or objective != 'pred_noise':
            raise NotImplementedError(f"safediffcon_b200.GaussianDiffusion: options outside the 1D hot path: {bad or objective}")
        if not (temporal and use_conv2d):
            raise NotImplementedError("safediffcon_b200.GaussianDiffusion covers the temporal 2-D (Nt, Nx) layout only")
        assert type(seq_length) is tuple and len(seq_length) == 2, \
            "should be a tuple of (Nt, Nx) (time evolution of a 1-d function)"
        self.model = model
        self.channels = self.model.channels
        self.self_condition = self.model.self_condition
        if self.self_condition:
            raise NotImplementedError("self-conditioning is not part of the 1D hot path")
        self.temporal, self.conv2d, self.traj_size = True, True, seq_length
        self.objective = objective

        if beta_schedule == 'linear':
            betas = linear_beta_schedule(timesteps)
        elif beta_schedule == 'cosine':
            betas = cosine_beta_schedule(timesteps)
        else:
            raise ValueError(f'unknown beta schedule {beta_schedule}')
        alphas = 1. - betas
        abar = torch.cumprod(alphas, dim=0)
        abar_prev = torch.cat([torch.ones(1, dtype=torch.float64), abar[:-1]])
        self.num_timesteps = int(betas.shape[0])
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else self.num_timesteps
        assert self.sampling_timesteps <= self.num_timesteps
        self.is_ddim_sampling = self.sampling_timesteps < self.num_timesteps
        self.ddim_sampling_eta = ddim_sampling_eta

        def reg(name, val):
            self.register_buffer(name, val.to(torch.float32))

        post_var = betas * (1. - abar_prev) / (1. - abar)
        reg('betas', betas)
        self.alphas = alphas.to(torch.float32).clone()
        self.alphas_prev = torch.cat([torch.ones(1, dtype=torch.float64), alphas[:-1]]).to(torch.float32)
        reg('alphas_cumprod', abar)
        reg('alphas_cumprod_prev', abar_prev)
        reg('sqrt_alphas_cumprod', torch.sqrt(abar))
        reg('sqrt_one_minus_alphas_cumprod', torch.sqrt(1. - abar))
        reg('log_one_minus_alphas_cumprod', torch.log(1. - abar))
        reg('sqrt_recip_alphas_cumprod', torch.sqrt(1. / abar))
        reg('sqrt_recipm1_alphas_cumprod', torch.sqrt(1. / abar - 1))
        reg('posterior_variance', post_var)
        reg('posterior_log_variance_clipped', torch.log(post_var.clamp(min=1e-20)))
        reg('posterior_mean_coef1', betas * torch.sqrt(abar_prev) / (1. - abar))
        reg('posterior_mean_coef2', (1. - abar_prev) * torch.sqrt(alphas) / (1. - abar))
        reg('loss_weight', torch.ones_like(abar))

        self.guidance_u0 = guidance_u0
        self.is_condition_u0 = is_condition_u0
        self.is_condition_uT = is_condition_uT
        self.is_condition_u0_zero_pred_noise = is_condition_u0_zero_pred_noise
        self.is_condition_uT_zero_pred_noise = is_condition_uT_zero_pred_noise
        self.train_on_partially_observed = train_on_partially_observed
        self.train_on_padded_locations = train_on_padded_locations
        self.condition_idx = condition_idx
        self.prior_beta = prior_beta
        self._tables = {}
        self._graphs = _GraphCache()
        # the denoiser's cached FiLM table covers integer times [0, table_timesteps): size it for this process
        if hasattr(self.model, "table_timesteps"):
            self.model.table_timesteps = max(int(self.model.table_timesteps), self.num_timesteps)

    def _load_from_state_dict(self, *args, **kwargs):
        # schedule buffers may change under a cached per-step table / captured graph: drop both
        self._tables = {}
        self._graphs = _GraphCache()
        return super()._load_from_state_dict(*args, **kwargs)

    # ------------------------------------------------------------------ helpers
    def predict_start_from_noise(self, x_t, t, noise):
        return extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - \
            extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise

    def predict_noise_from_start(self, x_t, t, x0):
        return (extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - x0) / \
            extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape)

    def q_posterior(self, x_start, x_t, t):
        mean = extract(self.posterior_mean_coef1, t, x_t.shape) * x_start + extract(self.posterior_mean_coef2, t, x_t.shape) * x_t
        return mean, extract(self.posterior_variance, t, x_t.shape), extract(self.posterior_log_variance_clipped, t, x_t.shape)

    def get_guidance_options(self, **kwargs):
        nabla_J = kwargs.get('nablaJ')
        if nabla_J is None:
            nabla_J = lambda x: 0  # noqa: E731
        sched = kwargs.get('J_scheduler') or (lambda t: 1.)
        proj = kwargs.get('proj_guidance') or (lambda ep, g: ep + g)
        return nabla_J, sched, proj

    def ddim_time_pairs(self):
        times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)
        times = list(reversed(times.int().tolist()))
        return list(zip(times[:-1], times[1:]))

    def _eps(self, img, t_int):
        """Denoiser output for a batch-uniform diffusion time (no host sync for the native U-Net)."""
        if hasattr(self.model, "denoise_uniform"):
            return self.model.denoise_uniform(img, t_int)
        tb = torch.full((img.shape[0],), t_int, device=img.device, dtype=torch.long)
        return self.model(img, tb, None, residual=None)

    def _coef_table(self, sampler, J_scheduler):
        """Per-step scalars, computed with the same fp32 torch-CPU ops the reference applies to its buffers."""
        # the schedule buffers are state-dict entries: load_state_dict / .to() may replace their contents, so their storage
        # address and version counter are part of the key (the reference always reads the live buffers)
        used = (self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod, self.alphas_cumprod, self.posterior_mean_coef1,
                self.posterior_mean_coef2, self.posterior_log_variance_clipped)
        key = (sampler, self.sampling_timesteps, float(self.ddim_sampling_eta), id(J_scheduler), str(self.betas.device),
               tuple((b.data_ptr(), b._version) for b in used))
        if J_scheduler is None and key in self._tables:
            return self._tables[key]
        c1b = self.sqrt_recip_alphas_cumprod.cpu()
        c2b = self.sqrt_recipm1_alphas_cumprod.cpu()
        rows, times = [], []
        sched = J_scheduler or (lambda t: 1.)
        if sampler == SAMPLER_DDIM:
            abar = self.alphas_cumprod.cpu()
            eta = self.ddim_sampling_eta
            for t, tn in self.ddim_time_pairs():
                if tn < 0:
                    rows.append((c1b[t].item(), c2b[t].item(), 0., 0., 0., float(sched(t)), 1, t))
                else:
                    a, an = abar[t], abar[tn]
                    sigma = eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
                    c = (1 - an - sigma ** 2).sqrt()
                    rows.append((c1b[t].item(), c2b[t].item(), an.sqrt().item(), c.item(), float(sigma), float(sched(t)), 0, t))
                times.append(t)
        else:
            m1, m2 = self.posterior_mean_coef1.cpu(), self.posterior_mean_coef2.cpu()
            lv = self.posterior_log_variance_clipped.cpu()
            for t in reversed(range(self.num_timesteps)):
                sd = (0.5 * lv[t]).exp().item()
                rows.append((c1b[t].item(), c2b[t].item(), m1[t].item(), m2[t].item(), sd, float(sched(t)), int(t == 0), t))
                times.append(t)
        arr = (L.StepCoef * len(rows))(*[L.StepCoef(*r) for r in rows])
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        dev = host.to(self.betas.device)
        tab = (dev, times, rows)
        if J_scheduler is None:
            self._tables[key] = tab
        return tab

    def set_condition(self, img, u, shape, u0_or_uT):
        if len(shape) != 4:
            raise ValueError('Bad sample shape')
        if u0_or_uT == 'uT':
            img[:, 0, self.condition_idx, :] = u
        elif u0_or_uT == 'u0':
            img[:, 0, 0, :] = u
        else:
            assert False

    def set_pad_condition(self, img, origin_img=None):
        if origin_img is None:
            origin_img = torch.zeros_like(img)
        img[..., 0, self.condition_idx + 1:, :] = origin_img[..., 0, self.condition_idx + 1:, :]
        img[..., 1, self.condition_idx:, :] = origin_img[..., 1, self.condition_idx:, :]
        img[..., 2, self.condition_idx:, :] = origin_img[..., 2, self.condition_idx:, :]

    # ------------------------------------------------------------------ fused step plumbing
    def _conditions(self, kwargs, w_groundtruth, device):
        u0 = L.dev_f32(kwargs['u_init'].to(device), 'u_init') if self.is_condition_u0 else None
        uT = L.dev_f32(kwargs['u_final'].to(device), 'u_final') if self.is_condition_uT else None
        wg = L.dev_f32(w_groundtruth.to(device), 'w_groundtruth') if w_groundtruth is not None else None
        return u0, uT, wg

    def _step(self, sampler, x, eps, noise, out, table, step, gstruct, grad, conds, clip_denoised, seed, offset,
              x0_out=None, eps_out=None, counter=None):
        B, C, H, W = x.shape
        pad = conds is not None and not self.train_on_padded_locations
        u0, uT, wg = conds if conds is not None else (None, None, None)
        L.check(L.lib().sdc_reverse_step(
            sampler, L.ptr(x), L.ptr(eps), L.ptr(noise), L.ptr(out), L.ptr(x0_out), L.ptr(eps_out), L.ptr(table), int(step),
            L.ptr(counter), ctypes.byref(gstruct) if gstruct is not None else None, L.ptr(grad), L.ptr(u0), L.ptr(uT), L.ptr(wg),
            self.condition_idx, int(pad), int(bool(clip_denoised)), seed, offset, B, H, W,
            L.stream_ptr()))

    def _initial(self, shape, device, noise_iter, seed, offset):
        if noise_iter is not None:
            return L.dev_f32(next(noise_iter).to(device), 'noise').clone()
        img = torch.empty(shape, device=device, dtype=torch.float32)
        L.check(L.lib().sdc_fill_normal(L.ptr(img), shape[0], img[0].numel(), seed, offset, _INIT_TAG, L.stream_ptr()))
        return img

    def _rng(self, kwargs, device):
        noise = kwargs.get('noise')
        noise_iter = iter(noise) if noise is not None else None
        seed = kwargs.get('seed')
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # follows torch.manual_seed like the reference's randn
        return noise_iter, int(seed), int(kwargs.get('sample_offset', 0))

    def _guidance_plan(self, kwargs):
        """-> (gstruct or None, generic nablaJ callable or None, proj callable or None)."""
        nablaJ = kwargs.get('nablaJ')
        proj = kwargs.get('proj_guidance')
        if not self.guidance_u0 or nablaJ is None:
            return None, None, None
        assert not self.self_condition, 'self condition not tested with guidance'
        if isinstance(nablaJ, SafetyGuidance) and proj is None:
            return nablaJ.struct(), None, None
        return None, nablaJ, proj

    def _guided_eps(self, x, eps, t_int, coef_row, clip, nablaJ, proj):
        """Generic (user-callable) guidance: x0 estimate -> nablaJ -> eps update, in torch ops on the device."""
        c1, c2, sched = coef_row[0], coef_row[1], coef_row[5]
        x0 = c1 * x - c2 * eps
        if clip:
            x0 = x0.clamp(-1., 1.)
        with torch.enable_grad():
            g = nablaJ(x0.clone().detach().requires_grad_())
        g = g * sched
        return (eps + g) if proj is None else proj(eps, g)

    # ------------------------------------------------------------------ captured-graph chain (small batches)
    def _graph_entry(self, sampler, shape, table, rows, gstruct, conds, has_noise, clip_denoised):
        """One reverse step (denoiser + fused posterior update in place + chain-state advance) captured in a CUDA graph
        over static buffers.  Step index / diffusion time, Philox seed and sample offset live in device memory
        (sdc_chain_state), and the denoiser's packed weights are refreshed in place when parameters change, so one graph
        serves every step of every chain with this shape, sampler and guidance."""
        B, C, H, W = shape
        device = self.betas.device
        # the denoiser's weights are (re)packed here, IN PLACE when they changed: every chain passes through this call
        plan = self.model._plan_ready() if hasattr(self.model, "_plan_ready") else None
        if plan is not None:
            wkey, pk = ("plan", id(plan), plan.prec), None
        else:
            pk = self.model._packed()
            self.model._film_table(pk)
            wkey = ("pack", id(pk), pk.get("table_gen", 0), pk["prec"])
        gkey = None if gstruct is None else tuple(getattr(gstruct, f) for f, _ in gstruct._fields_)
        key = (id(self.model), wkey, shape, sampler, tuple(rows), gkey, tuple(c is not None for c in conds), has_noise,
               self.condition_idx, bool(self.train_on_padded_locations), bool(clip_denoised))
        cache = self._graphs.entries
        if key in cache:
            cache.move_to_end(key)
            return cache[key]
        lib = L.lib()
        e = types.SimpleNamespace(pk=pk, plan=plan, table=table, n_steps=len(rows))
        e.ws = plan.workspace(B, H, W, device) if plan is not None else None   # the captured launches keep their workspace alive
        e.img = torch.zeros(shape, device=device)
        e.t_index = torch.zeros(B, dtype=torch.int32, device=device)
        e.state = torch.zeros(ctypes.sizeof(L.ChainState), dtype=torch.uint8, device=device)
        e.u0 = torch.zeros(B, W, device=device) if conds[0] is not None else None
        e.uT = torch.zeros(B, W, device=device) if conds[1] is not None else None
        e.wg = torch.zeros(B, H, W, device=device) if conds[2] is not None else None
        e.noise = torch.zeros(shape, device=device) if has_noise else None
        pad = int(not self.train_on_padded_locations)

        def one_step():
            eps = self.model.denoise_indexed(e.img, e.t_index, workspace=e.ws) if e.ws is not None else \
                self.model.denoise_indexed(e.img, e.t_index)
            L.check(lib.sdc_reverse_step_state(
                sampler, L.ptr(e.img), L.ptr(eps), L.ptr(e.noise), L.ptr(e.img), None, None, L.ptr(table), L.ptr(e.state),
                ctypes.byref(gstruct) if gstruct is not None else None, None, L.ptr(e.u0), L.ptr(e.uT), L.ptr(e.wg),
                self.condition_idx, pad, int(bool(clip_denoised)), B, H, W, L.stream_ptr()))
            L.check(lib.sdc_chain_state_advance(L.ptr(e.state), L.ptr(table), e.n_steps, L.ptr(e.t_index), B, L.stream_ptr()))

        L.check(lib.sdc_chain_state_set(L.ptr(e.state), 0, 0, 0, L.ptr(table), e.n_steps, L.ptr(e.t_index), B, L.stream_ptr()))
        one_step()   # eager warm-up on the static buffers: lazy initialisation happens outside the capture
        n0 = L.launch_count()
        e.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(e.graph):
            one_step()
        e.n_launch = L.launch_count() - n0
        cache[key] = e
        while len(cache) > GRAPH_CACHE_ENTRIES:
            cache.popitem(last=False)
        return e

    def _last_step_with_grad(self, img, t, coef_row, kwargs):
        """Final DDIM pair (t, -1) under autograd: x0 = clamp(c1 x - c2 eps'), eps' = eps + nablaJ(x0.detach()) * sched,
        the same fp32 mul / sub / clamp sequence as the fused kernel, on a denoiser output that carries parameter
        gradients (reference diffusion.py:524-531 with model_predictions :226-286)."""
        c1, c2, sched = coef_row[0], coef_row[1], coef_row[5]
        nablaJ, proj = kwargs.get('nablaJ'), kwargs.get('proj_guidance')
        with torch.enable_grad():
            tb = torch.full((img.shape[0],), t, device=img.device, dtype=torch.long)
            eps = self.model(img, tb)
            x0 = (c1 * img - c2 * eps).clamp(-1., 1.)
            if self.guidance_u0 and nablaJ is not None:
                g = nablaJ(x0.clone().detach().requires_grad_()) * sched
                eps = (eps + g) if proj is None else proj(eps, g)
                x0 = (c1 * img - c2 * eps).clamp(-1., 1.)
            return x0

    def _graph_chain(self, sampler, shape, table, times, rows, gstruct, conds, noise_iter, seed, offset, clip_denoised,
                     n_steps=None):
        B, C, H, W = shape
        lib = L.lib()
        e = self._graph_entry(sampler, tuple(shape), table, rows, gstruct, conds, noise_iter is not None, clip_denoised)
        table = e.table   # the captured launches read the entry's own copy (equal rows => equal contents)
        for dst, src in zip((e.u0, e.uT, e.wg), conds):
            if dst is not None:
                dst.copy_(src.reshape(dst.shape))
        if noise_iter is not None:
            e.img.copy_(next(noise_iter))
        else:
            L.check(lib.sdc_fill_normal(L.ptr(e.img), B, e.img[0].numel(), seed, offset, _INIT_TAG, L.stream_ptr()))
        L.check(lib.sdc_write_conditions(L.ptr(e.img), L.ptr(e.u0), L.ptr(e.uT), L.ptr(e.wg), self.condition_idx,
                                         int(not self.train_on_padded_locations), B, H, W, L.stream_ptr()))
        L.check(lib.sdc_chain_state_set(L.ptr(e.state), 0, seed & 0xFFFFFFFFFFFFFFFF, offset, L.ptr(table), e.n_steps,
                                        L.ptr(e.t_index), B, L.stream_ptr()))
        for step in range(len(times) if n_steps is None else n_steps):
            if noise_iter is not None and step != len(times) - 1:
                e.noise.copy_(next(noise_iter))
            e.graph.replay()
            lib.sdc_count_launches(e.n_launch)
        return e.img.clone()

    # ------------------------------------------------------------------ reference API
    def model_predictions(self, x, t, x_self_cond=None, residual=None, clip_x_start=False, rederive_pred_noise=False, **kwargs):
        """(pred_noise, pred_x_start) for per-sample times t (reference diffusion.py:226-286), torch ops on the device."""
        model_output = self.model(x, t, x_self_cond, residual=residual)
        clipf = (lambda v: v.clamp(-1., 1.)) if clip_x_start else (lambda v: v)
        nablaJ, sched, proj = self.get_guidance_options(**kwargs)
        pred_noise = kwargs['pred_noise'] if kwargs.get('pred_noise') is not None else model_output
        if kwargs.get('pred_noise') is not None:
            assert self.guidance_u0 is False, 'guidance should be w.r.t. ut'
        x_start = clipf(self.predict_start_from_noise(x, t, pred_noise))
        if self.guidance_u0:
            with torch.enable_grad():
                x_clone = x_start.clone().detach().requires_grad_()
                pred_noise = proj(pred_noise, nablaJ(x_clone) * sched(t[0].item()))
        x_start = clipf(self.predict_start_from_noise(x, t, pred_noise))
        if clip_x_start and rederive_pred_noise:
            pred_noise = self.predict_noise_from_start(x, t, x_start)
        return ModelPrediction(pred_noise, x_start)

    def p_mean_variance(self, x, t, x_self_cond=None, residual=None, **kwargs):
        """(model_mean, posterior_variance, posterior_log_variance, x_start, pred_noise) for per-sample times t (reference
        diffusion.py:288-297): model_predictions without inner clamps, in-place clamp of x_start, q_posterior.  Torch ops on
        the device around the CUDA denoiser (the chains use the fused sdc_reverse_step instead; parity: tests/test_chain_gpu.py)."""
        preds = self.model_predictions(x, t, x_self_cond, residual=residual, **kwargs)
        x_start = preds.pred_x_start
        if kwargs['clip_denoised']:   # a required key, as in the reference (KeyError when absent)
            x_start.clamp_(-1., 1.)
        model_mean, posterior_variance, posterior_log_variance = self.q_posterior(x_start=x_start, x_t=x, t=t)
        return model_mean, posterior_variance, posterior_log_variance, x_start, preds.pred_noise

    @torch.no_grad()
    def p_sample(self, x, t: int, x_self_cond=None, residual=None, **kwargs):
        """One DDPM step at integer time t -> (pred_img, x_start, pred_noise) (reference diffusion.py:299-306)."""
        x = L.dev_f32(x, 'x')
        table, times, rows = self._coef_table(SAMPLER_DDPM, kwargs.get('J_scheduler'))
        step = self.num_timesteps - 1 - t
        noise_iter, seed, offset = self._rng(kwargs, x.device)
        gstruct, nablaJ, proj = self._guidance_plan(kwargs)
        eps = kwargs['pred_noise'] if kwargs.get('pred_noise') is not None else self._eps(x, t)
        if kwargs.get('pred_noise') is not None:
            assert self.guidance_u0 is False, 'guidance should be w.r.t. ut'
        elif nablaJ is not None:
            eps = self._guided_eps(x, eps, t, rows[step], False, nablaJ, proj)
        eps = L.dev_f32(eps, 'eps')
        z = L.dev_f32(next(noise_iter).to(x.device), 'noise') if (noise_iter is not None and t > 0) else None
        out, x0, en = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        # conditions are NOT written by p_sample itself (the loop writes them before the next step)
        with torch.cuda.device(x.device):
            self._step(SAMPLER_DDPM, x, eps, z, out, table, step, gstruct, None, None,
                       kwargs.get('clip_denoised', True), seed, offset, x0, en)
        return out, x0, en

    @torch.no_grad()
    def p_sample_loop(self, shape, w_groundtruth=None, enable_grad=True, **kwargs):
        assert not self.is_ddim_sampling, 'wrong branch!'
        device = self._require_cuda()
        with torch.cuda.device(device):
            return self._run_chain(SAMPLER_DDPM, shape, w_groundtruth, enable_grad, False, kwargs)

    @torch.no_grad()
    def ddim_sample(self, shape, return_all_timesteps=False, w_groundtruth=None, enable_grad=False, **kwargs):
        device = self._require_cuda()
        with torch.cuda.device(device):
            return self._run_chain(SAMPLER_DDIM, shape, w_groundtruth, enable_grad, return_all_timesteps, kwargs)

    def _require_cuda(self):
        device = self.betas.device
        if device.type != 'cuda':
            raise RuntimeError("safediffcon_b200.GaussianDiffusion.sample: module is on the CPU; move it to a CUDA device "
                               "(there is no CPU fallback)")
        return device

    def _run_chain(self, sampler, shape, w_groundtruth, enable_grad, return_all, kwargs):
        """One sampling chain with the FP16-range guard: the FP16 executor counts non-finite eps entries on the device; the
        count is read ONCE after the chain.  If it is non-zero the denoiser switches itself to TF32 operands (same 10-bit
        mantissa, fp32 range), warns, and the chain is repeated with the same seed / noise -- a checkpoint whose activations
        leave +-65504 never returns silently corrupted samples.  Non-finite values that survive TF32 are data (the reference
        would produce them too) and are returned as they are."""
        net = self.model
        guard = hasattr(net, "take_nonfinite") and getattr(net, "precision", None) == "f16" and getattr(net, "overflow_fallback", True)
        if guard:
            if kwargs.get('noise') is not None and not isinstance(kwargs['noise'], (list, tuple)):
                kwargs = dict(kwargs, noise=list(kwargs['noise']))   # re-iterable for the repeat
            if kwargs.get('noise') is None and kwargs.get('seed') is None:
                kwargs = dict(kwargs, seed=int(torch.randint(0, 2 ** 62, (1,)).item()))   # same stream for the repeat
            net.take_nonfinite()
        out = self._run_chain_once(sampler, shape, w_groundtruth, enable_grad, return_all, kwargs)
        if guard and net.take_nonfinite() > 0:
            import warnings
            warnings.warn("safediffcon_b200: the FP16 denoiser produced non-finite values (activations beyond the fp16 range?); "
                          "switching Unet2D.precision to 'tf32' and repeating the chain", RuntimeWarning)
            net.precision = "tf32"
            out = self._run_chain_once(sampler, shape, w_groundtruth, enable_grad, return_all, kwargs)
        return out

    def _run_chain_once(self, sampler, shape, w_groundtruth, enable_grad, return_all, kwargs):
        device = self._require_cuda()
        if hasattr(self.model, "revalidate_packed"):
            # EMA / optimiser updates written through p.data do not bump version counters: compare the parameter digest once
            # per chain (one small sync next to 200-1000 denoiser evaluations) and repack when it moved
            self.model.revalidate_packed()
        table, times, rows = self._coef_table(sampler, kwargs.get('J_scheduler'))
        noise_iter, seed, offset = self._rng(kwargs, device)
        conds = self._conditions(kwargs, w_groundtruth, device)
        gstruct, nablaJ, proj = self._guidance_plan(kwargs)
        clip_denoised = kwargs.get('clip_denoised', True)
        ddim = sampler == SAMPLER_DDIM
        B, C, H, W = shape

        second_call = (not ddim) and (not self.guidance_u0)  # DDPM calibration-style path: two p_sample calls per step
        # the reference runs the LAST DDIM step under torch.enable_grad() (diffusion.py:524-551): when the denoiser has
        # trainable parameters the returned x0 then carries the autograd graph InferenceFT.finetune_step back-propagates
        with torch.enable_grad():
            grad_last = bool(ddim and enable_grad and not return_all and hasattr(self.model, "wants_param_grad")
                             and self.model.wants_param_grad())
        n_plain = len(times) - (1 if grad_last else 0)
        if (0 < B <= GRAPH_MAX_BATCH and hasattr(self.model, "denoise_indexed") and nablaJ is None and not second_call
                and not return_all and max(times) < getattr(self.model, "table_timesteps", 0)):
            img = self._graph_chain(sampler, shape, table, times, rows, gstruct, conds, noise_iter, seed, offset, clip_denoised,
                                    n_plain)
            return self._last_step_with_grad(img, times[-1], rows[-1], kwargs) if grad_last else img
        img = self._initial(shape, device, noise_iter, seed, offset)
        L.check(L.lib().sdc_write_conditions(L.ptr(img), L.ptr(conds[0]), L.ptr(conds[1]), L.ptr(conds[2]), self.condition_idx,
                                             int(not self.train_on_padded_locations), B, H, W, L.stream_ptr()))
        imgs = [img.clone()] if return_all else None
        nxt = torch.empty_like(img)
        for step, t in enumerate(times):
            last = step == len(times) - 1
            if last and grad_last:
                return self._last_step_with_grad(img, t, rows[step], kwargs)
            eps = L.dev_f32(self._eps(img, t), 'eps')
            z = None
            if noise_iter is not None and not last:
                if second_call:
                    z1 = next(noise_iter)  # draw consumed by the first p_sample
                z = L.dev_f32(next(noise_iter).to(device), 'noise')
            if second_call:
                if last and enable_grad:
                    break  # reference quirk: the t=0 branch keeps img when guidance_u0 is False (diffusion.py:445-447)
                user_nabla = kwargs.get('nablaJ')
                if user_nabla is not None:
                    # img_curr from the first p_sample feeds the guidance, eps' feeds the second (diffusion.py:416-423)
                    cur = torch.empty_like(img)
                    z1d = L.dev_f32(z1.to(device), 'noise') if (noise_iter is not None and not last) else None
                    self._step(sampler, img, eps, z1d, cur, table, step, None, None, None, clip_denoised,
                               seed ^ 0x5DEECE66D, offset)
                    sched = rows[step][5]
                    g = user_nabla(cur) * sched
                    eps = L.dev_f32(eps + g if kwargs.get('proj_guidance') is None else kwargs['proj_guidance'](eps, g), 'eps')
                self._step(sampler, img, eps, z, nxt, table, step, None, None, conds, clip_denoised, seed, offset)
            else:
                if nablaJ is not None:
                    eps = L.dev_f32(self._guided_eps(img, eps, t, rows[step], ddim, nablaJ, proj), 'eps')
                self._step(sampler, img, eps, z, nxt, table, step, gstruct, None, conds, clip_denoised, seed, offset)
            img, nxt = nxt, img
            if return_all:
                imgs.append(img.clone())
        return img if not return_all else torch.stack(imgs, dim=1)

    @torch.no_grad()
    def sample(self, batch_size=16, clip_denoised=True, w_groundtruth=None, enable_grad=True, **kwargs):
        """Same keyword contract as the reference (diffusion.py:557-607): nablaJ, J_scheduler, proj_guidance,
        guidance_u0, u_init, u_final, w_groundtruth; unknown keys (device, w_scheduler, timesteps, ...) are ignored.
        Extra keys of this implementation: ``noise`` (iterable of pre-generated draws, in the reference's draw
        order), ``seed`` / ``sample_offset`` (in-kernel Philox stream; sample_offset = global index of sample 0)."""
        if 'guidance_u0' in kwargs:
            self.guidance_u0 = kwargs['guidance_u0']
        if self.is_condition_u0:
            assert 'is_condition_u0' not in kwargs, 'specify this value in the model. not during sampling.'
            assert 'u_init' in kwargs and kwargs['u_init'] is not None
        if self.is_condition_uT:
            assert 'is_condition_uT' not in kwargs, 'specify this value in the model. not during sampling.'
            assert 'u_final' in kwargs and kwargs['u_final'] is not None
        sample_size = (batch_size, self.channels, *self.traj_size)
        sample_fn = self.p_sample_loop if not self.is_ddim_sampling else self.ddim_sample
        return sample_fn(sample_size, clip_denoised=clip_denoised, w_groundtruth=w_groundtruth, enable_grad=enable_grad, **kwargs)

    # ------------------------------------------------------------------ training loss
    def q_sample(self, x_start, t, noise=None):
        """x_t ~ q(x_t | x_0) (reference diffusion.py:629-636)."""
        noise = torch.randn_like(x_start) if noise is None else noise
        return extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start + \
            extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise

    def p_losses(self, x_start, t, noise=None, mean=True):
        """Diffusion training loss of the reference (diffusion.py:638-733) for the options this implementation covers
        (pred_noise, u0/uT conditioning, pad masking).  The denoiser call is the CUDA forward whose backward yields the
        parameter gradients (unet.py: _TrainFn); the surrounding [B,3,16,128] algebra is plain torch, as in the reference.
        The reference zeroes the conditioned rows of the caller's `noise` tensor in place; so does this."""
        noise = torch.randn_like(x_start) if noise is None else noise
        x = self.q_sample(x_start=x_start, t=t, noise=noise)
        if self.is_condition_u0:
            self.set_condition(x, x_start[:, 0, 0, :], x.shape, 'u0')
        if self.is_condition_uT:
            self.set_condition(x, x_start[:, 0, self.condition_idx, :], x.shape, 'uT')
        if not self.train_on_padded_locations:
            self.set_pad_condition(x)
        model_out = self.model(x, t, None, residual=None)
        target = noise
        if self.train_on_partially_observed is not None:
            raise NotImplementedError("train_on_partially_observed is outside the 1D hot path")
        if self.is_condition_u0 and self.is_condition_u0_zero_pred_noise:
            self.set_condition(noise, torch.zeros_like(x[:, 0, 0, :]), x.shape, 'u0')
        if self.is_condition_uT and self.is_condition_uT_zero_pred_noise:
            self.set_condition(noise, torch.zeros_like(x[:, 0, 0, :]), x.shape, 'uT')
        if not self.train_on_padded_locations:
            model_out = model_out.clone()
            self.set_pad_condition(model_out, origin_img=target)
        loss = torch.nn.functional.mse_loss(model_out, target, reduction='none')
        loss = loss.reshape(loss.shape[0], -1).mean(dim=1)
        loss = loss * extract(self.loss_weight, t, loss.shape)
        return loss.mean() if mean else loss

    def forward(self, img, *args, **kwargs):
        """loss = p_losses(img, t ~ U{0..T-1}) (reference diffusion.py:735-746)."""
        b, c, nt, nx = img.shape
        assert (nt, nx) == tuple(self.traj_size), f'traj size must be (nt, nx) of ({nt, nx}), but got {self.traj_size}'
        t = torch.randint(0, self.num_timesteps, (b,), device=img.device).long()
        return self.p_losses(img, t, *args, **kwargs)
