"""Drop-in Python surface of the reference's class-conditional DDPM sampling path, executed by libsgb200.

Mirrors /root/reference/src/diff_modules.py:
  * UNet_conditional(c_in, c_out, time_dim, num_classes, remove_deep_conv).forward(x, t, y)   (:204-217)
    -- same constructor, same 183-key state_dict (names and shapes), same call signature;
  * Diffusion(noise_steps, beta_start, beta_end, img_size, num_classes, c_in, c_out, device, **kwargs)
    with .beta/.alpha/.alpha_hat/.model/.img_size/.device/.c_in/.num_classes, .prepare_noise_schedule(),
    .sample(use_ema, labels, cfg_scale=3) -> uint8 [n, c_in, S, S]  (:411-442) and .load(dir, ...) (:509-510);
    the upstream positional form sample(model, n, labels, cfg_scale) is accepted as an alias;
  * EMA.reset_parameters (:48-49) -- the load side of EMA; the reference never builds `ema_model`
    (:393 is commented out, so sample(use_ema=True) raises AttributeError there); here load() fills it
    from `ema_ckpt.pt` when that file exists.

The modules hold parameters only.  All arithmetic runs in hand-written sm_100a kernels behind the C ABI
(include/sgb200.h); there is no torch-op or CPU fallback, and a non-B200 device is an error.
"""
from __future__ import annotations

import copy
import math
import os

import torch
import torch.nn as nn

from . import _cabi, ops
from .engine import MODES, TIME_DIM, PackedWeights, PlanOptions, UNetPlan
from .vae import VqaeDecoder

__all__ = ["UNet_conditional", "Diffusion", "DiffusionVAE", "EMA", "state_dict_schema"]


# ----------------------------------------------------------------------------------------------------
# state_dict schema (names/shapes of the reference's nn.Module tree)
# ----------------------------------------------------------------------------------------------------
def _dc(p, cin, cout, cmid=None):
    cmid = cmid or cout
    return [(f"{p}.double_conv.0.weight", (cmid, cin, 3, 3)), (f"{p}.double_conv.1.weight", (cmid,)),
            (f"{p}.double_conv.1.bias", (cmid,)), (f"{p}.double_conv.3.weight", (cout, cmid, 3, 3)),
            (f"{p}.double_conv.4.weight", (cout,)), (f"{p}.double_conv.4.bias", (cout,))]


def _emb(p, cout):
    return [(f"{p}.emb_layer.1.weight", (cout, TIME_DIM)), (f"{p}.emb_layer.1.bias", (cout,))]


def _sa(p, c):
    return [(f"{p}.mha.in_proj_weight", (3 * c, c)), (f"{p}.mha.in_proj_bias", (3 * c,)),
            (f"{p}.mha.out_proj.weight", (c, c)), (f"{p}.mha.out_proj.bias", (c,)),
            (f"{p}.ln.weight", (c,)), (f"{p}.ln.bias", (c,)),
            (f"{p}.ff_self.0.weight", (c,)), (f"{p}.ff_self.0.bias", (c,)),
            (f"{p}.ff_self.1.weight", (c, c)), (f"{p}.ff_self.1.bias", (c,)),
            (f"{p}.ff_self.3.weight", (c, c)), (f"{p}.ff_self.3.bias", (c,))]


def state_dict_schema(c_in=1, c_out=1, num_classes=None, remove_deep_conv=False):
    """Ordered (key, shape) list, identical to the reference model's state_dict."""
    s = _dc("inc", c_in, 64)
    for name, cin, cout, sa_c in (("down1", 64, 128, 128), ("down2", 128, 256, 256), ("down3", 256, 256, 256)):
        s += _dc(f"{name}.maxpool_conv.1", cin, cin) + _dc(f"{name}.maxpool_conv.2", cin, cout) + _emb(name, cout)
        s += _sa(f"sa{name[-1]}", sa_c)
    if remove_deep_conv:
        s += _dc("bot1", 256, 256) + _dc("bot3", 256, 256)
    else:
        s += _dc("bot1", 256, 512) + _dc("bot2", 512, 512) + _dc("bot3", 512, 256)
    for i, (name, cin, cout) in enumerate((("up1", 512, 128), ("up2", 256, 64), ("up3", 128, 64))):
        s += _dc(f"{name}.conv.0", cin, cin) + _dc(f"{name}.conv.1", cin, cout, cin // 2) + _emb(name, cout)
        s += _sa(f"sa{4 + i}", cout)
    s += [("outc.weight", (c_out, 64, 1, 1)), ("outc.bias", (c_out,))]
    if num_classes is not None:
        s += [("label_emb.weight", (num_classes, TIME_DIM))]
    return s


class _Params(nn.Module):
    """A bare container node; the tree of these reproduces the reference's parameter names."""


def _init_param(name, shape):
    """torch's default initialisers for the corresponding reference layers."""
    t = torch.empty(shape)
    leaf = name.rsplit(".", 1)[-1]
    is_norm = any(s in name for s in (".double_conv.1.", ".double_conv.4.", ".ln.", ".ff_self.0."))
    if name == "label_emb.weight":
        return nn.init.normal_(t)
    if is_norm:
        return nn.init.ones_(t) if leaf == "weight" else nn.init.zeros_(t)
    if name.endswith("in_proj_weight"):
        return nn.init.xavier_uniform_(t)
    if name.endswith("in_proj_bias") or name.endswith("out_proj.bias"):
        return nn.init.zeros_(t)
    if leaf == "weight":
        return nn.init.kaiming_uniform_(t, a=math.sqrt(5))
    return t  # Linear / outc biases: filled by the caller from the weight's fan-in


class UNet_conditional(nn.Module):
    """ε-prediction UNet with timestep and class conditioning (reference :139-217), B200-only."""

    def __init__(self, c_in=1, c_out=1, time_dim=256, num_classes=None, remove_deep_conv=False,
                 compute_dtype="fp32", **kwargs):
        super().__init__()
        if time_dim != TIME_DIM:
            # the reference hard-wires emb_dim=256 in Down/Up (:97,:117), so any other time_dim fails there too
            raise ValueError("time_dim must be 256 (Down/Up hard-wire emb_dim=256 in the reference)")
        if kwargs:
            raise TypeError(f"unexpected arguments {sorted(kwargs)}")
        self.c_in, self.c_out = c_in, c_out
        self.time_dim = time_dim
        self.num_classes = num_classes
        self.remove_deep_conv = remove_deep_conv
        self.compute_dtype = compute_dtype
        fan = {}
        for key, shape in state_dict_schema(c_in, c_out, num_classes, remove_deep_conv):
            *path, leaf = key.split(".")
            node = self
            for part in path:
                if part not in node._modules:
                    node.add_module(part, _Params())
                node = node._modules[part]
            t = _init_param(key, shape)
            if leaf == "weight" and len(shape) >= 2:
                fan[".".join(path)] = int(torch.tensor(shape[1:]).prod())
            elif leaf == "bias" and ".".join(path) in fan and not any(
                    s in key for s in (".double_conv.", ".ln.", ".ff_self.0.", "out_proj")):
                bound = 1.0 / math.sqrt(fan[".".join(path)])
                nn.init.uniform_(t, -bound, bound)
            node.register_parameter(leaf, nn.Parameter(t))
        self._packed = None
        self._packed_key = None
        self._plans = {}
        self._weights_gen = 0  # bumped by writers that bypass autograd's version counter (EMA.update_model_average)
        self._raw16_ok = True  # cleared when a GroupNorm input left fp16's safe range: plans then keep raw tensors in fp32

    # ------------------------------------------------------------------ engine plumbing
    def set_compute_dtype(self, mode: str):
        """'fp32' (split-TF32 tensor-core engine, <= 1e-4 of the reference), 'bf16' / 'f16' (tcgen05 16-bit engine) or
        'fp32_simt' (CUDA-core kernels: the comparator of the parity tests)."""
        if mode not in MODES:
            raise ValueError(f"compute dtype must be one of {sorted(MODES)}")
        self.compute_dtype = mode
        return self

    def _weights_key(self):
        ps = list(self.parameters())
        return (self.compute_dtype, str(ps[0].device), self._weights_gen, tuple(p._version for p in ps),
                tuple(p.data_ptr() for p in ps))

    def invalidate_packed(self):
        """Drop the kernel-layout weight copy and every plan / captured graph built on it.  Writers that change parameter
        storage in place through raw pointers (sg_ema_update) call this: such writes bump neither Parameter._version nor
        data_ptr, so the cache key alone would not notice them."""
        self._weights_gen += 1
        self._packed = None
        self._packed_key = None
        self._plans = {}

    def packed_weights(self) -> PackedWeights:
        """Kernel-layout weights; rebuilt whenever a parameter was modified (load_state_dict, .to(), ...)."""
        key = self._weights_key()
        if self._packed is None or key != self._packed_key:
            dev = next(self.parameters()).device
            _cabi.require_b200(dev)
            self._packed = PackedWeights(self.state_dict(), dev, self.compute_dtype)
            self._packed_key = key
            self._plans = {}
        return self._packed

    def plan(self, *, n_src, rows, S, use_step=False, debug=False, options: PlanOptions | None = None) -> UNetPlan:
        w = self.packed_weights()
        if options is None:
            options = PlanOptions(raw16=self._raw16_ok)
        key = (n_src, rows, S, use_step, debug, options)
        if key not in self._plans:
            self._plans[key] = UNetPlan(w, n_src=n_src, rows=rows, S=S, use_step=use_step, debug=debug, options=options)
        return self._plans[key]

    def fp16_range_fallback(self, plan) -> bool:
        """GroupNorm(1, C) is scale invariant in the reference; the fp16 raw conv outputs of the 16-bit engines are not.
        After a run: if a GroupNorm input left fp16's safe range (the kernels compare the exact fp32 statistics against
        [2^-10, 2^10] rms and raise a flag), drop every plan and switch this model to fp32 raw tensors.  Returns True
        when the caller has to repeat the run."""
        if not plan.range_overflow():
            return False
        self._raw16_ok = False
        self._plans = {}
        return True

    def release_plans(self):
        self._plans = {}

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def forward(self, x, t, y=None):
        """x [n, c_in, S, S]; t [n] (long or float); y [n] int64 class ids or None.  Returns ε [n, c_out, S, S] fp32."""
        if x.dim() != 4 or x.shape[1] != self.c_in or x.shape[2] != x.shape[3]:
            raise ValueError(f"x must be [n, {self.c_in}, S, S], got {tuple(x.shape)}")
        if y is not None and self.num_classes is None:
            raise AttributeError("'UNet_conditional' object has no attribute 'label_emb'")  # as the reference would
        n, S = x.shape[0], x.shape[2]
        if y is not None and n > 0:
            y = torch.as_tensor(y).reshape(-1)
            if len(y) != n:
                raise ValueError(f"y must have one class id per sample ({n}), got {len(y)}")
            if int(y.min()) < 0 or int(y.max()) >= self.num_classes:
                raise IndexError("index out of range in self")  # nn.Embedding's error in the reference (:215)
        while True:
            plan = self.plan(n_src=n, rows=n, S=S)
            plan.reset_range_flag()
            plan.x_in.copy_(x.to(torch.float32))
            plan.t.copy_(t.reshape(-1).to(torch.float32))
            if y is None:
                plan.y.fill_(-1)
            else:
                plan.y.copy_(y.reshape(-1).to(torch.int64))
            plan.run()
            if not self.fp16_range_fallback(plan):
                return plan.eps.clone()


class EMA:
    """The reference's EMA helper (:24-49); the average itself is one kernel launch per parameter tensor."""

    def __init__(self, beta=0.995):
        self.beta = beta
        self.step = 0

    def update_model_average(self, ma_model, current_model):
        for current_params, ma_params in zip(current_model.parameters(), ma_model.parameters()):
            ma_params.data = self.update_average(ma_params.data, current_params.data)
        if hasattr(ma_model, "invalidate_packed"):
            ma_model.invalidate_packed()  # the packed copy / plans / graphs of ma_model are stale now

    def update_average(self, old, new):
        """old * beta + (1 - beta) * new (:37-40), in place on `old` (the reference rebinds .data to a new tensor)."""
        if old is None:
            return new
        return ops.ema_update(old, new.to(old.device), self.beta)

    def step_ema(self, ema_model, model, step_start_ema=2000):
        if self.step < step_start_ema:
            self.reset_parameters(ema_model, model)
            self.step += 1
            return
        self.update_model_average(ema_model, model)
        self.step += 1

    def reset_parameters(self, ema_model, model):
        ema_model.load_state_dict(model.state_dict())


def _viridis():
    """matplotlib.cm.viridis, imported lazily (the reference imports matplotlib at module top)."""
    try:
        from matplotlib import cm
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise ImportError("the PNG writers colour with matplotlib.cm.viridis (as the reference does); install "
                          "matplotlib or pass colormap=") from e
    return cm.viridis


class Diffusion:
    """Linear-β DDPM with classifier-free-guidance ancestral sampling (reference :370-442)."""

    def __init__(self, noise_steps=1000, beta_start=1e-4, beta_end=0.02, img_size=256, num_classes=10, c_in=1,
                 c_out=1, device="cuda", **kwargs):
        self.noise_steps = noise_steps
        self.beta_start = beta_start
        self.beta_end = beta_end
        self.device = torch.device(device)
        _cabi.require_b200(self.device)
        # (:387-389) evaluated on the host exactly like the CPU reference, then moved
        beta = self.prepare_noise_schedule()
        alpha = 1.0 - beta
        alpha_hat = torch.cumprod(alpha, dim=0)
        self.beta, self.alpha, self.alpha_hat = beta.to(self.device), alpha.to(self.device), alpha_hat.to(self.device)
        # posterior coefficients with the reference's own expressions (:436-439)
        self._coef = torch.stack([1 / torch.sqrt(alpha), (1 - alpha) / (torch.sqrt(1 - alpha_hat)), torch.sqrt(beta)],
                                 dim=1).contiguous().to(self.device)
        self.img_size = img_size
        self.model = UNet_conditional(c_in, c_out, num_classes=num_classes, **kwargs).to(self.device)
        self.ema_model = None  # filled by load() when an EMA checkpoint exists
        self.c_in = c_in
        self.num_classes = num_classes
        self.gpu_launches = 0  # kernels launched by the last sample() call
        self.graph_captures = 0  # CUDA graphs captured so far (a repeated sample() call replays the cached one)

    def prepare_noise_schedule(self):
        return torch.linspace(self.beta_start, self.beta_end, self.noise_steps)

    # ------------------------------------------------------------------ forward-only training helpers (:401-409, :474-478)
    def sample_timesteps(self, n):
        return torch.randint(low=1, high=self.noise_steps, size=(n,))

    def noise_images(self, x, t, *, noise=None, seed=0, sample_base=0):
        """Reference `noise_images(x, t)` -> (x_t, eps): x_t = sqrt(alpha_hat[t]) x + sqrt(1 - alpha_hat[t]) eps.
        `noise` (ours) injects eps (then x_t is bit-identical to the reference's expression); otherwise eps comes from
        the Philox stream keyed by (seed, sample_base + sample index)."""
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        t = torch.as_tensor(t).reshape(-1).to(device=self.device, dtype=torch.int64)
        if noise is not None:
            noise = noise.to(device=self.device, dtype=torch.float32).contiguous()
        return ops.noise_images(x, t, self.alpha_hat, eps=noise, seed=seed, sample_base=sample_base)

    def mse(self, noise, pred):
        """The reference's `self.mse = nn.MSELoss()` (:392, used at :478)."""
        return ops.mse(noise.to(self.device, torch.float32).contiguous(), pred.to(self.device, torch.float32).contiguous())

    @torch.no_grad()
    def eval_loss(self, images, labels=None, *, t=None, noise=None, seed=0):
        """The body of `one_epoch(train=False)` for one batch (:474-478): t ~ sample_timesteps, noise_images, model
        forward (labels=None is the unconditional pass the training loop takes 10 % of the time), MSE(noise, pred)."""
        n = images.shape[0]
        t = self.sample_timesteps(n) if t is None else t
        x_t, eps = self.noise_images(images, t, noise=noise, seed=seed)
        pred = self.model(x_t, torch.as_tensor(t).to(self.device), labels)
        return self.mse(eps, pred)

    # ------------------------------------------------------------------ checkpoints (:509-510)
    def load(self, model_cpkt_path, model_ckpt="ckpt.pt", ema_model_ckpt="ema_ckpt.pt"):
        self.model.load_state_dict(torch.load(os.path.join(model_cpkt_path, model_ckpt), weights_only=True))
        ema_path = os.path.join(model_cpkt_path, ema_model_ckpt)
        if os.path.exists(ema_path):
            self.ema_model = copy.deepcopy(self.model).eval().requires_grad_(False)
            self.ema_model.load_state_dict(torch.load(ema_path, weights_only=True))

    def prepare(self, args):
        """Reference `prepare(args)` (:548-560) as far as sampling needs it: the run folders of `mk_folders`
        (models/<run_name>, results/<run_name>) and the EMA helper.  The data loaders, AdamW, OneCycleLR and GradScaler
        the reference also builds here belong to training (no backward pass in this package)."""
        os.makedirs(os.path.join("models", args.run_name), exist_ok=True)
        os.makedirs(os.path.join("results", args.run_name), exist_ok=True)
        self.ema = EMA(0.995)

    def load_model(self, args):
        """Reference `load_model(args)` (:525-546): loads models/<run_name>/ckpt.pt when args.load_model is set and
        raises FileNotFoundError when a checkpoint is missing.  The reference also restores optim.pt into its AdamW
        optimizer; sampling has no optimizer, so the file is only required to exist (same error as upstream)."""
        if not args.load_model:
            print("Starting model ftesh...")
            return
        model_path = os.path.join("models", args.run_name, "ckpt.pt")
        if os.path.exists(model_path):
            self.model.load_state_dict(torch.load(model_path, weights_only=True))
            print(f"Model loaded successfully from {model_path}")
        else:
            raise FileNotFoundError(f"Model checkpoint not found at {model_path}")
        optim_path = os.path.join("models", args.run_name, "optim.pt")
        if not os.path.exists(optim_path):
            raise FileNotFoundError(f"Optimizer checkpoint not found at {optim_path}")
        print("Model loaded successfully")

    # ------------------------------------------------------------------ generation driver (:759-775)
    def gen_images(self, img_folder, samp_i, labels=None, *, colormap=None, **sample_kw):
        """Reference `gen_images(img_folder, samp_i, labels=None)`: one sample per label, coloured with viridis and
        written as RGBA PNG `{class_name}_gen_imgs_{i}_{samp_i}.png` (the name format src/helpers.py:602-610 parses).
        `colormap` (ours): callable uint8 [H, W] -> float RGBA [H, W, 4]; default matplotlib.cm.viridis, imported
        lazily (the reference imports matplotlib at module top).  Returns the list of files written."""
        return self.gen_images_many(img_folder, [samp_i], labels, colormap=colormap, **sample_kw)

    def gen_images_many(self, img_folder, samp_is, labels=None, *, colormap=None, **sample_kw):
        """`gen_images` for several `samp_i` in ONE sampling call (ours): the label set is tiled len(samp_is) times so
        that the batch fills the GPU (27 labels alone run at ~75 % of the throughput of a 512-sample batch).  Files and
        pixels are those of the per-`samp_i` calls when the caller passes `sample_base = samp_is[0] * len(labels)`
        and consecutive `samp_is` (the Philox stream is keyed by the global sample index)."""
        if labels is None:
            labels = torch.arange(self.num_classes).long().to(self.device)
        labels = torch.as_tensor(labels).reshape(-1)
        samp_is = list(samp_is)
        if getattr(self, "sav_denoise_path", None):
            # (:765-769) the trajectory dumps are named by class only, so the label set is sampled ONCE (tiling it would
            # rewrite every {class}_noise_{i}_*.png len(samp_is) times) and the final images are not saved
            self.sample(False, labels, **sample_kw)
            print("not saving image, just noise portions")
            return []
        sampled_images = self.sample(False, labels.repeat(len(samp_is)), **sample_kw)
        return self.write_images(img_folder, samp_is, labels, sampled_images, colormap=colormap)

    def write_images(self, img_folder, samp_is, labels, sampled_images, *, colormap=None, workers=None):
        """Host half of gen_images (:766-775): colour map + PNG files for images [len(samp_is) * len(labels), 1, H, W]
        (device or host tensor; the copy to the host happens here, so a caller may run this on a worker thread while the
        next batch is being sampled).  The PNG encoder (zlib, ~20 ms per 256x256 RGBA image) releases the GIL, so the
        files are written by `workers` threads (default: the host cores this rank can claim, at most 8)."""
        import numpy as np
        from concurrent.futures import ThreadPoolExecutor
        from PIL import Image

        if colormap is None:
            colormap = getattr(self, "colormap", None) or _viridis()
        class_names = getattr(self, "class_names", None) or [str(k) for k in range(self.num_classes or 0)]
        lab_list = torch.as_tensor(labels).reshape(-1).tolist() if samp_is is not None else None
        host = sampled_images.cpu().numpy()
        jobs = []
        if samp_is is None:
            # `labels` = explicit (class id, i, samp_i) triples, one per image: a slice of the flattened (samp_i, class) list
            for k, (lab, i, samp_i) in enumerate(labels):
                jobs.append((k, f"{img_folder}/{class_names[lab]}_gen_imgs_{i}_{samp_i}.png"))
        else:
            for g, samp_i in enumerate(samp_is):
                for i, lab in enumerate(lab_list):
                    jobs.append((g * len(lab_list) + i, f"{img_folder}/{class_names[lab]}_gen_imgs_{i}_{samp_i}.png"))

        def one(job):
            k, path = job
            rgba = colormap(host[k].transpose(1, 2, 0).squeeze())
            rgba = (np.asarray(rgba) * 255).astype(np.uint8)
            Image.fromarray(rgba).save(path)
            return path

        if workers is None:
            world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
            workers = max(1, min(8, (os.cpu_count() or 1) // max(world, 1)))
        if workers == 1 or len(jobs) < 2:
            return [one(j) for j in jobs]
        with ThreadPoolExecutor(max_workers=workers) as ex:
            return list(ex.map(one, jobs))

    # ------------------------------------------------------------------ sampling (:411-442)
    @torch.no_grad()
    def sample(self, use_ema, labels, cfg_scale=3, *legacy, noise=None, seed=0, sample_base=0, micro_batch=512,
               return_float=False, use_graph=True, max_steps=None, step_hook=None, return_trajectory=False):
        """Reference form: sample(use_ema, labels, cfg_scale=3) -> uint8 [n, c_in, S, S].
        Upstream alias: sample(model, n, labels, cfg_scale=3).

        Extra keyword-only controls (ours):
          noise        fp32 [T-1, n, c, S, S] injected Gaussians: noise[0] = x_T, noise[k] = z at i = T-k
                       (the order the reference consumes its generator).  Default: Philox4x32-10 keyed by
                       (seed, sample_base + sample index, timestep) -- independent of batch split / GPU count.
          micro_batch  samples per captured loop (n is processed in chunks of this size)
          return_float return the fp32 state before the uint8 quantisation
          max_steps    run only the first k loop iterations (benchmarking a bounded number of timesteps)
          return_trajectory  also return the fp32 states [K+1, n, c, S, S]: x_T, then x after each of the K executed
                       loop iterations (K = T-1 unless max_steps) -> (out, trajectory)
          step_hook    callable(i, x, labels) run on the host after the update of loop index i (T-1 ... 1) with the
                       chunk's fp32 state x (read-only) -- the captured step is replayed once per timestep, so the hook
                       sits between two replays and costs nothing when absent
        """
        if isinstance(use_ema, nn.Module):  # upstream (model, n, labels, cfg_scale)
            model, n_expected = use_ema, labels
            labels = cfg_scale
            cfg_scale = legacy[0] if legacy else 3
            if len(labels) != n_expected:
                raise ValueError("n does not match len(labels)")
        else:
            if legacy:
                raise TypeError("too many positional arguments")
            if use_ema and self.ema_model is None:
                raise AttributeError("'Diffusion' object has no attribute 'ema_model' (no EMA checkpoint loaded)")
            model = self.ema_model if use_ema else self.model
        labels = torch.as_tensor(labels).reshape(-1).to(device=self.device, dtype=torch.int64)
        n = len(labels)
        S, c, T = self.img_size, self.c_in, self.noise_steps
        if n > 0 and self.num_classes is not None and (int(labels.min()) < 0 or int(labels.max()) >= self.num_classes):
            raise IndexError("index out of range in self")  # nn.Embedding's error in the reference
        if noise is not None:
            noise = noise.to(device=self.device, dtype=torch.float32)
            if tuple(noise.shape) != (T - 1, n, c, S, S):
                raise ValueError(f"noise must be [T-1, n, c, S, S] = {(T - 1, n, c, S, S)}, got {tuple(noise.shape)}")
        out_dtype = torch.float32 if return_float else torch.uint8
        out = torch.empty((n, c, S, S), dtype=out_dtype, device=self.device)
        self.gpu_launches = 0
        traj = None
        if return_trajectory:
            iters = T - 1 if max_steps is None else min(T - 1, max_steps)
            traj = torch.empty((iters + 1, n, c, S, S), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            for lo in range(0, n, micro_batch):
                hi = min(n, lo + micro_batch)
                nz = None if noise is None else noise[:, lo:hi].contiguous()
                args = (model, labels[lo:hi], float(cfg_scale), nz, seed, sample_base + lo, out[lo:hi], use_graph,
                        max_steps, step_hook, None if traj is None else traj[:, lo:hi])
                if self._sample_chunk(*args):  # fp16 range guard tripped: the model now plans with fp32 raw tensors
                    self._sample_chunk(*args)
        return (out, traj) if return_trajectory else out

    def _sample_chunk(self, model, labels, cfg, noise, seed, sample_base, out, use_graph, max_steps, step_hook=None,
                      traj=None):
        """One micro-batch through the captured loop.  Returns True when the run has to be repeated (fp16 range guard)."""
        n = len(labels)
        T = self.noise_steps
        rows = 2 * n if cfg > 0 else n
        plan = model.plan(n_src=n, rows=rows, S=self.img_size, use_step=True)
        x = plan.x_in  # the sampler state lives in the plan's input buffer: no copy per step
        plan.y.fill_(-1)
        plan.y[:n].copy_(labels)

        def one_step():
            plan.run()
            ops.cfg_update(x, plan.eps, self._coef, plan.step, cfg_scale=cfg, noise=noise, seed=seed,
                           sample_base=sample_base)
            ops.step_advance(plan.step)

        launches_per_step = plan.n_launches + 2
        if not getattr(plan, "_warm", False):
            plan.step.fill_(T - 1)
            one_step()  # loads the kernels' modules and sets their attributes outside of graph capture
            torch.cuda.synchronize(self.device)
            plan._warm = True
            self.gpu_launches += launches_per_step
        plan.reset_range_flag()
        # x_T (:418) and the step counter i = T-1
        if noise is not None:
            x.copy_(noise[0])
        else:
            ops.philox_normal(x, seed=seed, sample_base=sample_base, step_tag=T)
            self.gpu_launches += 1
        plan.step.fill_(T - 1)
        iters = T - 1 if max_steps is None else min(T - 1, max_steps)
        if traj is not None:
            traj[0].copy_(x)
        if use_graph and iters > 0:
            # the captured step bakes in its by-value launch arguments, so the graph is cached with the plan (whose
            # buffers it references) under exactly those: guidance scale and Philox key.  Injected-noise runs (parity
            # tests) capture afresh: their graph reads a caller-owned buffer that must not be kept alive here.
            gkey = (cfg, int(seed), int(sample_base), id(self._coef))
            graphs = plan.__dict__.setdefault("_graphs", {})
            entry = graphs.get(gkey) if noise is None else None
            if entry is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    one_step()
                # capture does not execute: state is still (x_T, T-1)
                entry = (g,)
                if noise is None:
                    if len(graphs) >= 8:
                        graphs.pop(next(iter(graphs)))
                    graphs[gkey] = entry
                self.graph_captures += 1
            step = entry[0].replay
        else:
            step = one_step
        for k in range(iters):
            step()
            if traj is not None:
                traj[k + 1].copy_(x)
            if step_hook is not None:
                step_hook(T - 1 - k, x, labels)
        self.gpu_launches += iters * launches_per_step
        if model.fp16_range_fallback(plan):
            return True
        if out.dtype == torch.uint8:
            ops.to_uint8(x, out)
            self.gpu_launches += 1
        else:
            out.copy_(x)
        return False


class DiffusionVAE(Diffusion):
    """Latent-space variant (reference :578-706): the UNet denoises [n, 4, S/4, S/4] latents and the VQAE codebook +
    decoder turn them into [n, 1, S, S] spectrograms.  Same constructor as the reference; `vqae_path` is loaded with
    torch.load(weights_only=True) unless `vqae_state_dict` (ours, keyword-only) is given.  Only the decode side of
    the VQAE is on the sampling path.  `sav_denoise_path` (:661-700) writes, at loop indices i % 50 == 0, i == 1 and
    i == T-1, `{class}_noise_{i}_latent.png` (the quantised latent's four channels as a 2x2 grid) and
    `{class}_noise_{i}_decode.png` (the decoded spectrogram) per sample, both viridis RGBA like the reference;
    `colormap` (ours, keyword-only) replaces matplotlib.cm.viridis (callable: uint8 or float-in-[0,1] array -> RGBA
    floats)."""

    def __init__(self, noise_steps=1000, beta_start=1e-4, beta_end=0.02, img_size=256, num_classes=10, c_in=1, c_out=1,
                 device="cuda", vqae_path="models/VQAE/ckpt.pt", sav_denoise_path=None, class_names=(), *,
                 vqae_state_dict=None, colormap=None, **kwargs):
        latent_dim = 4  # (:611); the reference builds UNet_conditional(latent_dim, latent_dim) at img_size // 4 (:624-627)
        super().__init__(noise_steps, beta_start, beta_end, img_size // 4, num_classes, latent_dim, latent_dim, device,
                         **kwargs)
        if vqae_state_dict is None:
            vqae_state_dict = torch.load(vqae_path, map_location="cpu", weights_only=True)
        # the decode tail is 0.08 % of the sampling FLOPs and has no normalisation behind its convolutions: in the fp32
        # modes it runs on the CUDA-core kernels (bit-level agreement with the reference's uint8 image), not on split TF32
        mode = self.model.compute_dtype
        self.vqae = VqaeDecoder(vqae_state_dict, self.device, "fp32_simt" if mode.startswith("fp32") else mode)
        self.class_names = list(class_names)
        self.sav_denoise_path = sav_denoise_path
        self.colormap = colormap
        self.dump_launches = 0  # kernels launched by the trajectory dumps of the last sample() call

    def dump_steps(self):
        """Loop indices at which the reference dumps the trajectory (:662)."""
        T = self.noise_steps
        return [i for i in range(T - 1, 0, -1) if i % 50 == 0 or i == 1 or i == T - 1]

    def _dump_denoise(self, i, x, labels):
        """The body of `if self.sav_denoise_path:` (:661-700) for the chunk state x [n, 4, S, S]."""
        if not (i % 50 == 0 or i == 1 or i == self.noise_steps - 1):
            return
        import numpy as np
        from PIL import Image

        cmap = self.colormap if self.colormap is not None else _viridis()
        print(f"saving denoise at step {i}...")
        up, q = self.vqae.decode(x, return_quantized=True)
        lat = ops.to_uint8_wrap(q)
        self.dump_launches += self.vqae.gpu_launches + 1
        lat, up = lat.cpu().numpy(), up.cpu().numpy()
        for img, img_up, lab in zip(lat, up, labels.tolist()):
            grid = np.concatenate([np.concatenate([img[0], img[1]], axis=1),
                                   np.concatenate([img[2], img[3]], axis=1)], axis=0)
            rgba = (np.asarray(cmap(grid / 255.0)) * 255).astype(np.uint8)  # float input, as the reference passes it
            Image.fromarray(rgba).save(f"{self.sav_denoise_path}/{self.class_names[lab]}_noise_{i}_latent.png")
            rgba = (np.asarray(cmap(img_up[0])) * 255).astype(np.uint8)     # uint8 input: a direct LUT index
            Image.fromarray(rgba).save(f"{self.sav_denoise_path}/{self.class_names[lab]}_noise_{i}_decode.png")

    @torch.no_grad()
    def sample(self, use_ema, labels, cfg_scale=3, *legacy, decode_micro_batch=64, **kw):
        """sample(use_ema, labels, cfg_scale=3) -> uint8 [n, 1, 4*img_size, 4*img_size] (:630-706); with
        return_trajectory=True -> (images, latent trajectory fp32 [K+1, n, 4, img_size, img_size])."""
        if kw.pop("return_float", False):
            raise TypeError("DiffusionVAE.sample returns the decoded uint8 image; use Diffusion.sample for latents")
        self.dump_launches = 0
        if self.sav_denoise_path and kw.get("step_hook") is None:
            kw["step_hook"] = self._dump_denoise
        x = super().sample(use_ema, labels, cfg_scale, *legacy, return_float=True, **kw)
        traj = None
        if kw.get("return_trajectory"):
            x, traj = x  # latent trajectory [K+1, n, 4, S/4, S/4]
        launches = self.gpu_launches + self.dump_launches
        out = self.vqae.decode(x, micro_batch=decode_micro_batch)
        self.gpu_launches = launches + self.vqae.gpu_launches
        return (out, traj) if traj is not None else out
