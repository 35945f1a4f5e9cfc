"""End-to-end parity of the B200 path with the reference, through the drop-in Python surface
(UNet_conditional.forward / Diffusion.sample) and therefore through the C ABI.

Comparators:
  * tests/golden/golden.npz -- outputs of the UNMODIFIED reference (tests/golden/make_golden.py);
  * oracle/ddpm_oracle.py    -- the CPU restatement (itself pinned to the golden vectors), for inputs the
                                fixtures do not cover.
Tolerances (BASELINE.json north_star): per-step eps rel-L2 <= 1e-4 in fp32 mode, <= 1e-2 in the 16-bit
tensor-core modes; 50-step trajectories |dx| <= 1e-3 (fp32) / <= 0.1 and <= 8 uint8 levels (16-bit).
"""
import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle.weights import make_state_dict
from tests.golden.make_golden import EPS_CASES, NUM_CLASSES, TRAJ_CASES, WEIGHT_SEED, golden_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"
# "fp32" = the fp32-accurate engine on tensor cores (split-TF32), "fp32_simt" = the CUDA-core comparator
EPS_TOL = {"fp32": 1e-4, "fp32_simt": 1e-4, "bf16": 1e-2, "f16": 1e-2}
MODES = ["fp32", "fp32_simt", "bf16", "f16"]


def build_model(mode, c=4, remove_deep_conv=False, num_classes=NUM_CLASSES):
    from spectrogramgenai_b200.diff_modules import UNet_conditional

    m = UNet_conditional(c, c, num_classes=num_classes, remove_deep_conv=remove_deep_conv, compute_dtype=mode)
    m.load_state_dict(make_state_dict(WEIGHT_SEED, c, c, num_classes, remove_deep_conv), strict=True)
    return m.to(DEV)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", EPS_CASES, ids=[c[0] for c in EPS_CASES])
def test_eps_matches_reference(golden, case, mode):
    tag, s, c, n, ts = case
    m = build_model(mode, c)
    x, y = golden_inputs(s, c, n)
    for tv in ts:
        t = (torch.ones(n) * tv).long()
        e_c = m(x.to(DEV), t.to(DEV), y.to(DEV)).cpu()
        e_u = m(x.to(DEV), t.to(DEV), None).cpu()
        assert e_c.shape == (n, c, s, s) and e_c.dtype == torch.float32
        err_c = O.rel_l2(e_c, torch.from_numpy(golden[f"eps_{tag}_t{tv}_cond"]))
        err_u = O.rel_l2(e_u, torch.from_numpy(golden[f"eps_{tag}_t{tv}_uncond"]))
        print(f"eps rel-L2 {tag} t={tv} {mode}: cond {err_c:.3e} uncond {err_u:.3e}")
        assert err_c < EPS_TOL[mode] and err_u < EPS_TOL[mode]


@pytest.mark.parametrize("mode", MODES)
def test_per_block_taps(golden, mode):
    """Every block output against the reference's forward hooks (R16, t=500, one sample)."""
    m = build_model(mode)
    x, y = golden_inputs(16, 4, 2)
    plan = m.plan(n_src=1, rows=1, S=16, debug=True)
    plan.x_in.copy_(x[:1])
    plan.t.fill_(500.0)
    plan.y.copy_(y[:1])
    plan.run()
    torch.cuda.synchronize()
    tol = 2e-5 if mode.startswith("fp32") else 1e-2
    for k in ["inc", "down1", "sa1", "down2", "sa2", "down3", "sa3", "bot1", "bot2", "bot3", "up1", "sa4", "up2",
              "sa5", "up3", "sa6"]:
        got = plan.taps[k].permute(0, 3, 1, 2).float().cpu()
        err = O.rel_l2(got, torch.from_numpy(golden[f"tap_r16_{k}"]))
        print(f"tap {k} {mode}: {err:.3e}")
        assert err < tol, k


@pytest.mark.parametrize("mode", ["fp32", "fp32_simt", "bf16"])
def test_shallow_variant(golden, mode):
    m = build_model(mode, remove_deep_conv=True)
    x, y = golden_inputs(16, 4, 2)
    e = m(x.to(DEV), (torch.ones(2) * 500).long().to(DEV), y.to(DEV)).cpu()
    assert O.rel_l2(e, torch.from_numpy(golden["eps_r16shallow_t500_cond"])) < EPS_TOL[mode]


@pytest.mark.parametrize("mode", ["fp32", "fp32_simt"])
def test_forward_accepts_float_t_and_odd_batches(mode):
    """t as float == t as long; batch sizes that do not fill a tile; batch invariance."""
    m = build_model(mode)
    sd = make_state_dict(WEIGHT_SEED, 4, 4, NUM_CLASSES)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(5, 4, 16, 16, generator=g)
    y = torch.randint(0, NUM_CLASSES, (5,), generator=g)
    t = torch.tensor([999, 3, 250, 1, 77])
    e_long = m(x.to(DEV), t.to(DEV), y.to(DEV)).cpu()
    e_float = m(x.to(DEV), t.float().to(DEV), y.to(DEV)).cpu()
    assert torch.equal(e_long, e_float)
    assert O.rel_l2(e_long, O.unet_forward(sd, x, t, y)) < 1e-4
    e_one = m(x[2:3].to(DEV), t[2:3].to(DEV), y[2:3].to(DEV)).cpu()
    assert torch.equal(e_one[0], e_long[2])  # per-sample results do not depend on the batch they ride in


@pytest.mark.parametrize("mode", ["bf16", "f16"])
def test_tensor_core_modes_small_and_ragged_batches(mode):
    """n = 1 and n = 5 at 16x16 latents: the deepest SelfAttention sees 64 .. 320 tokens, i.e. fewer than one
    128-token tile and a ragged last tile in the fused head / tail kernels, the attention core and the GEMMs."""
    m = build_model(mode)
    sd = make_state_dict(WEIGHT_SEED, 4, 4, NUM_CLASSES)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(5, 4, 16, 16, generator=g)
    y = torch.randint(0, NUM_CLASSES, (5,), generator=g)
    t = torch.tensor([999, 3, 250, 1, 77])
    want = O.unet_forward(sd, x, t, y)
    e5 = m(x.to(DEV), t.to(DEV), y.to(DEV)).cpu()
    assert O.rel_l2(e5, want) < EPS_TOL[mode]
    e1 = m(x[2:3].to(DEV), t[2:3].to(DEV), y[2:3].to(DEV)).cpu()
    assert O.rel_l2(e1, want[2:3]) < EPS_TOL[mode]
    assert torch.equal(e1[0], e5[2])  # per-sample results do not depend on the batch they ride in


def _diffusion(mode, s, T, c=4):
    from spectrogramgenai_b200.diff_modules import Diffusion

    d = Diffusion(noise_steps=T, img_size=s, num_classes=NUM_CLASSES, c_in=c, c_out=c, device=DEV, compute_dtype=mode)
    d.model.load_state_dict(make_state_dict(WEIGHT_SEED, c, c, NUM_CLASSES), strict=True)
    return d


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", TRAJ_CASES, ids=[c[0] for c in TRAJ_CASES])
def test_trajectory_matches_reference(golden, case, mode):
    """T=50 CFG sampling with the noise the reference drew under set_seed(7), CUDA-graph loop."""
    tag, s, c, n, T = case
    d = _diffusion(mode, s, T, c)
    y = torch.from_numpy(golden[f"traj_{tag}_T{T}_labels"])
    noise = O.draw_reference_noise(7, n, c, s, T)
    xf = d.sample(False, y, cfg_scale=3, noise=noise, return_float=True).cpu()
    ref = torch.from_numpy(golden[f"traj_{tag}_T{T}_xfloat"])
    dmax = float((xf - ref).abs().max())
    u8 = d.sample(False, y, cfg_scale=3, noise=noise).cpu()
    assert u8.dtype == torch.uint8 and u8.shape == (n, c, s, s)
    du8 = int(np.abs(u8.numpy().astype(int) - golden[f"traj_{tag}_T{T}_u8"].astype(int)).max())
    print(f"trajectory {tag} T={T} {mode}: max|dx| {dmax:.3e}, max uint8 diff {du8}, launches {d.gpu_launches}")
    if mode.startswith("fp32"):
        assert dmax < 1e-3 and du8 <= 1
    else:
        assert dmax < 0.1 and du8 <= 8
    # the quantised output is exactly the reference's tail applied to our float state
    assert torch.equal(u8, O.to_uint8(xf))


def test_cfg0_branch_and_eager_equals_graph(golden):
    d = _diffusion("fp32", 16, 12)
    _, y = golden_inputs(16, 4, 2)
    noise = O.draw_reference_noise(7, 2, 4, 16, 12)
    u8 = d.sample(False, y, cfg_scale=0, noise=noise).cpu().numpy().astype(int)
    assert np.abs(u8 - golden["traj_r16_T12_cfg0_u8"].astype(int)).max() <= 1
    a = d.sample(False, y, cfg_scale=3, noise=noise, return_float=True, use_graph=True)
    b = d.sample(False, y, cfg_scale=3, noise=noise, return_float=True, use_graph=False)
    assert torch.equal(a, b)
    # upstream positional alias sample(model, n, labels, cfg_scale)
    c = d.sample(d.model, 2, y, 3, noise=noise, return_float=True)
    assert torch.equal(a, c)


def test_philox_sampling_is_split_invariant_and_seeded():
    """Property test at the reference's configured latent size (R64): the result of sampling k spectrograms does
    not depend on how they are batched (micro-batches / ranks), and does depend on the seed."""
    d = _diffusion("bf16", 64, 6)
    y = torch.arange(6) % NUM_CLASSES
    a = d.sample(False, y, seed=11, return_float=True)
    b = d.sample(False, y, seed=11, return_float=True, micro_batch=4)
    c = torch.cat([d.sample(False, y[:3], seed=11, return_float=True, sample_base=0),
                   d.sample(False, y[3:], seed=11, return_float=True, sample_base=3)])
    assert torch.equal(a, b) and torch.equal(a, c)
    e = d.sample(False, y, seed=12, return_float=True)
    assert not torch.equal(a, e)
    assert torch.isfinite(a).all()


def test_ema_checkpoint_loading(tmp_path):
    """load(dir): ckpt.pt -> model, ema_ckpt.pt -> ema_model (the reference names the file and ignores it, :509-510)."""
    d = _diffusion("fp32", 16, 4)
    sd = make_state_dict(WEIGHT_SEED, 4, 4, NUM_CLASSES)
    sd_ema = make_state_dict(WEIGHT_SEED + 1, 4, 4, NUM_CLASSES)
    torch.save(sd, tmp_path / "ckpt.pt")
    y = torch.tensor([1, 2])
    with pytest.raises(AttributeError):
        d.sample(True, y)
    d.load(str(tmp_path))
    assert d.ema_model is None
    torch.save(sd_ema, tmp_path / "ema_ckpt.pt")
    d.load(str(tmp_path))
    noise = O.draw_reference_noise(3, 2, 4, 16, 4)
    got = d.sample(True, y, noise=noise, return_float=True).cpu()
    want = O.sample(sd_ema, y, noise, noise_steps=4, return_float=True)
    assert float((got - want).abs().max()) < 1e-3
    got_model = d.sample(False, y, noise=noise, return_float=True).cpu()
    assert not torch.allclose(got, got_model)
    for k, v in d.model.state_dict().items():
        assert torch.equal(v.cpu(), sd[k])


def test_bad_labels_raise():
    d = _diffusion("fp32", 16, 4)
    with pytest.raises(IndexError):
        d.sample(False, torch.tensor([0, NUM_CLASSES]))



@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_shared_prefix_is_bit_identical(mode):
    """With the conditional and unconditional halves batched (rows = 2n) the label-independent prefix (inc and the
    convolutions of down1) runs once for n rows; eps must be bit-identical to the plan that runs it for all 2n rows."""
    m = build_model(mode)
    n, S = 3, 32
    x, y = golden_inputs(S, 4, n)
    eps = {}
    from spectrogramgenai_b200.engine import PlanOptions

    for flag in ("1", "0"):
        m.release_plans()
        plan = m.plan(n_src=n, rows=2 * n, S=S, use_step=True, options=PlanOptions(shared_prefix=flag == "1"))
        assert plan.rows_p == (n if flag == "1" else 2 * n)
        plan.x_in.copy_(x.to(DEV))
        plan.y.fill_(-1)
        plan.y[:n].copy_(y.to(DEV))
        plan.step.fill_(321)
        plan.run()
        torch.cuda.synchronize()
        eps[flag] = plan.eps.clone()
    m.release_plans()
    assert torch.equal(eps["1"], eps["0"])
    assert not torch.equal(eps["1"][:n], eps["1"][n:])


def test_return_trajectory_matches_oracle_step_by_step():
    """return_trajectory: [K+1, n, c, S, S] = x_T and the state after every loop iteration, against the CPU oracle's
    trajectory with the same injected noise (fp32 engine, T = 12 at R16), also through micro-batches and max_steps."""
    T, s, c, n = 12, 16, 4, 3
    d = _diffusion("fp32", s, T, c)
    _, y = golden_inputs(s, c, n)
    noise = O.draw_reference_noise(7, n, c, s, T)
    ref = []
    O.sample(make_state_dict(WEIGHT_SEED, c, c, NUM_CLASSES), y, noise, noise_steps=T, trajectory=ref)
    u8, traj = d.sample(False, y, cfg_scale=3, noise=noise, return_trajectory=True)
    assert traj.shape == (T, n, c, s, s) and traj.dtype == torch.float32
    assert torch.equal(traj[0].cpu(), noise[0])
    for k, (i, _, x_ref) in enumerate(ref):
        assert i == T - 1 - k
        assert float((traj[k + 1].cpu() - x_ref).abs().max()) < 1e-3
    assert torch.equal(u8.cpu(), O.to_uint8(traj[-1].cpu()))
    u8b, trajb = d.sample(False, y, cfg_scale=3, noise=noise, return_trajectory=True, micro_batch=2, max_steps=5)
    assert trajb.shape == (6, n, c, s, s) and torch.equal(trajb, traj[:6])


def test_full_bench_batch_is_batch_invariant():
    """BASELINE configs[2] geometry at full size: one CFG step of n = 512 spectrograms at [4, 64, 64] (1024 UNet rows:
    32768 query tiles per attention head, persistent GEMM / fused-kernel loops tens of tiles deep, the shared
    label-independent prefix).  The oracle cannot run this size in seconds, so the check is the size-independent
    property: a sample's conditional and unconditional predictions are bit-identical to those it gets in a batch of 2
    (which the small-batch tests pin to the reference)."""
    from spectrogramgenai_b200.diff_modules import UNet_conditional

    m = UNet_conditional(4, 4, num_classes=NUM_CLASSES, compute_dtype="bf16")
    m.load_state_dict(make_state_dict(WEIGHT_SEED, 4, 4, NUM_CLASSES))
    m = m.to(DEV)
    n, S = 512, 64
    g = torch.Generator(device="cpu").manual_seed(9)
    x = torch.randn(n, 4, S, S, generator=g).to(DEV)
    y = (torch.arange(n) % NUM_CLASSES).to(DEV)
    plan = m.plan(n_src=n, rows=2 * n, S=S)
    plan.x_in.copy_(x)
    plan.t.fill_(437.0)
    plan.y.fill_(-1)
    plan.y[:n].copy_(y)
    plan.run()
    torch.cuda.synchronize()
    big = plan.eps.clone()
    assert torch.isfinite(big).all()
    m.release_plans()
    pick = [0, 255, 511]
    for i in pick:
        small = m.plan(n_src=2, rows=4, S=S)
        j = (i + 7) % n
        small.x_in.copy_(x[[i, j]])
        small.t.fill_(437.0)
        small.y.fill_(-1)
        small.y[:2].copy_(y[[i, j]])
        small.run()
        torch.cuda.synchronize()
        assert torch.equal(small.eps[0], big[i]) and torch.equal(small.eps[2], big[n + i]), i
        assert torch.equal(small.eps[1], big[j]) and torch.equal(small.eps[3], big[n + j]), j
    m.release_plans()


# ---------------------------------------------------------------------------------------------------------------------
# 256x256-spectrogram geometry (BASELINE configs[4]; Diffusion's default img_size=256, c_in=1, reference :376-378)
# ---------------------------------------------------------------------------------------------------------------------
LARGE_TOL = {"fp32": 1e-4, "bf16": 1e-2, "f16": 1e-2}


@pytest.mark.parametrize("mode", ["fp32", "bf16", "f16"])
@pytest.mark.parametrize("tag,s", [("p128c1", 128), ("p256c1", 256)])
def test_eps_matches_reference_at_256x256_geometry(golden_large, tag, s, mode):
    """eps at S = 128 (sa6: L = 16384) against the UNMODIFIED reference and at S = 256 (sa6: L = 65536, 512 key tiles per
    query tile, 2048 query tiles per head and sample) against the query-chunked oracle (tests/golden/make_golden_large.py)."""
    c, n, tv = 1, 1, 400
    m = build_model(mode, c)
    x, y = golden_inputs(s, c, n)
    t = (torch.ones(n) * tv).long()
    e_c = m(x.to(DEV), t.to(DEV), y.to(DEV)).cpu()
    err = O.rel_l2(e_c, torch.from_numpy(golden_large[f"eps_{tag}_t{tv}_cond"]))
    print(f"eps rel-L2 {tag} t={tv} {mode}: cond {err:.3e}")
    assert err < LARGE_TOL[mode]
    if f"eps_{tag}_t{tv}_uncond" in golden_large:
        e_u = m(x.to(DEV), t.to(DEV), None).cpu()
        err_u = O.rel_l2(e_u, torch.from_numpy(golden_large[f"eps_{tag}_t{tv}_uncond"]))
        print(f"eps rel-L2 {tag} t={tv} {mode}: uncond {err_u:.3e}")
        assert err_u < LARGE_TOL[mode]
    m.release_plans()


@pytest.mark.parametrize("mode", ["fp32", "bf16", "f16"])
def test_eps_matches_oracle_at_batch_64_r64(mode):
    """The latency sweep's geometry (BASELINE configs[1]): one forward of n = 64 samples at [4, 64, 64] against the CPU
    oracle on the same inputs (the oracle needs ~10 s for it), so that large-batch runs have a direct comparator and not
    only the batch-invariance property."""
    n, s, c = 64, 64, 4
    m = build_model(mode, c)
    x, y = golden_inputs(s, c, n, seed=321)
    t = torch.randint(1, 1000, (n,), generator=torch.Generator().manual_seed(5))
    want = O.unet_forward(make_state_dict(WEIGHT_SEED, c, c, NUM_CLASSES), x, t, y)
    got = m(x.to(DEV), t.to(DEV), y.to(DEV)).cpu()
    per_sample = ((got - want).flatten(1).norm(dim=1) / want.flatten(1).norm(dim=1))
    print(f"eps rel-L2 n=64 R64 {mode}: all {O.rel_l2(got, want):.3e}, worst sample {float(per_sample.max()):.3e}")
    assert O.rel_l2(got, want) < EPS_TOL[mode]
    assert float(per_sample.max()) < 1.5 * EPS_TOL[mode]
    m.release_plans()


# ---------------------------------------------------------------------------------------------------------------------
# GroupNorm(1, C) makes the reference invariant to the scale of every conv weight; fp16 raw tensors are not
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["bf16", "f16"])
@pytest.mark.parametrize("scale", [1e2, 1e4, 1e-3, 1e-5])
def test_conv_weight_scale_invariance(mode, scale):
    """Every 3x3 conv weight times `scale` (the oracle runs on the scaled weights too): the reference result is unchanged up to rounding because each conv feeds GroupNorm.  The 16-bit
    engines keep raw conv outputs in fp16 while they lie in fp16's safe range and otherwise fall back to fp32 raw tensors
    (range flag raised by GroupNorm-apply from the exact fp32 statistics) -- either way eps must stay within the bar."""
    n, s, c = 2, 16, 4
    sd = make_state_dict(WEIGHT_SEED, c, c, NUM_CLASSES)
    sd = {k: (v * scale if (v.dim() == 4 and v.shape[-1] == 3) else v) for k, v in sd.items()}
    from spectrogramgenai_b200.diff_modules import UNet_conditional

    m = UNet_conditional(c, c, num_classes=NUM_CLASSES, compute_dtype=mode)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    x, y = golden_inputs(s, c, n)
    t = torch.tensor([700, 30])
    want = O.unet_forward(sd, x, t, y)
    if scale > 1:  # (below 1 GroupNorm's eps = 1e-5 is no longer negligible against the variance: the function changes)
        base = O.unet_forward(make_state_dict(WEIGHT_SEED, c, c, NUM_CLASSES), x, t, y)
        assert O.rel_l2(want, base) < 1e-4  # the reference itself does not care about the scale
    got = m(x.to(DEV), t.to(DEV), y.to(DEV)).cpu()
    err = O.rel_l2(got, want)
    print(f"scale {scale:g} {mode}: eps rel-L2 {err:.3e}, fp16 raw tensors kept: {m._raw16_ok}")
    assert err < EPS_TOL[mode]
    if scale in (1e4, 1e-5):
        assert not m._raw16_ok  # far outside fp16's range: the guard must have switched to fp32 raw tensors
    # the sampler takes the same fallback
    from spectrogramgenai_b200.diff_modules import Diffusion

    d = Diffusion(noise_steps=6, img_size=s, num_classes=NUM_CLASSES, c_in=c, c_out=c, device=DEV, compute_dtype=mode)
    d.model.load_state_dict(sd, strict=True)
    noise = O.draw_reference_noise(3, n, c, s, 6)
    xs = d.sample(False, y, noise=noise, return_float=True).cpu()
    ref = O.sample(sd, y, noise, noise_steps=6, return_float=True)
    assert float((xs - ref).abs().max()) < 0.1


def test_ema_update_invalidates_packed_weights_and_plans():
    """EMA.update_model_average writes ma_model's parameters in place through raw pointers (sg_ema_update): neither
    Parameter._version nor data_ptr changes, so the packed-weight cache must be invalidated explicitly -- an ema_model that
    was already sampled from has to use the new average afterwards."""
    from spectrogramgenai_b200.diff_modules import EMA, UNet_conditional

    c = 4
    model, ema_model = build_model("fp32", c), build_model("fp32", c)
    model.load_state_dict(make_state_dict(WEIGHT_SEED + 7, c, c, NUM_CLASSES))
    x, y = golden_inputs(16, c, 2)
    t = torch.tensor([500, 20])
    before = ema_model(x.to(DEV), t.to(DEV), y.to(DEV)).clone()  # packs weights, builds a plan
    ema = EMA(0.5)
    ema.step = 10
    ema.step_ema(ema_model, model, step_start_ema=0)  # in-place average (beta = 0.5)
    after = ema_model(x.to(DEV), t.to(DEV), y.to(DEV)).clone()
    fresh = UNet_conditional(c, c, num_classes=NUM_CLASSES, compute_dtype="fp32").to(DEV)
    fresh.load_state_dict(ema_model.state_dict())
    want = fresh(x.to(DEV), t.to(DEV), y.to(DEV))
    assert not torch.equal(before, after)
    assert torch.equal(after, want)
    sd_avg = {k: v.cpu() for k, v in ema_model.state_dict().items()}
    assert O.rel_l2(after.cpu(), O.unet_forward(sd_avg, x, t, y)) < 1e-4


def test_captured_graph_is_cached_with_the_plan():
    """Diffusion.sample replays the SAME captured graph when called again with the same Philox key / guidance scale."""
    d = _diffusion("bf16", 16, 6)
    y = torch.tensor([1, 2, 3])
    a = d.sample(False, y, seed=5, return_float=True)
    n_cap = d.graph_captures
    b = d.sample(False, y, seed=5, return_float=True)
    assert d.graph_captures == n_cap and torch.equal(a, b)
    c = d.sample(False, y, seed=6, return_float=True)
    assert d.graph_captures == n_cap + 1 and not torch.equal(a, c)


def test_two_models_on_two_streams_do_not_share_state():
    """model and a second model (different weights) run concurrently on two streams: the small by-value weights of
    inc.double_conv.0 and of the fused output conv belong to each launch (no __constant__ bank rewritten per call), so the
    interleaved results equal the serial ones."""
    m1, m2 = build_model("bf16"), build_model("bf16")
    m2.load_state_dict(make_state_dict(WEIGHT_SEED + 3, 4, 4, NUM_CLASSES))
    x, y = golden_inputs(32, 4, 3)
    t = torch.tensor([900.0, 400.0, 10.0])
    plans = []
    for m in (m1, m2):
        p = m.plan(n_src=3, rows=3, S=32)
        p.x_in.copy_(x.to(DEV))
        p.t.copy_(t.to(DEV))
        p.y.copy_(y.to(DEV))
        plans.append(p)
    want = []
    for p in plans:
        p.run()
        torch.cuda.synchronize()
        want.append(p.eps.clone())
        p.eps.fill_(float("nan"))
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(6):  # no host synchronisation in between: the launches of the two plans interleave
        for p, st in zip(plans, streams):
            with torch.cuda.stream(st):
                p.run()
    torch.cuda.synchronize()
    assert torch.equal(plans[0].eps, want[0]) and torch.equal(plans[1].eps, want[1])
    assert not torch.equal(want[0], want[1])
