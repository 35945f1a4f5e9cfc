"""Generation driver: the B200 counterpart of /root/reference/src/ddpm_conditional_generate.py (same flags, same
output files).  One process per GPU; with torchrun the flattened list of (samp_i, class) spectrograms is split across
ranks in contiguous, balanced slices, so no collective is needed: every rank writes its own PNG files.

    python -m spectrogramgenai_b200.generate --run_name DDPM_conditional_VAE --num_samples 50 --img_folder diffusion_samples
    torchrun --nproc-per-node 8 -m spectrogramgenai_b200.generate ...
"""
from __future__ import annotations

import argparse
import json
import os
import time
from types import SimpleNamespace

import torch

from .diff_modules import DiffusionVAE
from .diff_utils import set_seed


def parse_args(argv=None):
    c = SimpleNamespace(run_name="DDPM_conditional_VAE", seed=42, num_classes=27, noise_steps=1000, img_size=256,
                        dataset_path="Birdnet_conf_files_images_split", img_folder="diffusion_samples", train_folder="train",
                        start_idx=0, num_samples=50, sav_denoise_path=None)
    p = argparse.ArgumentParser(description="CFG-DDPM spectrogram generation on B200")
    p.add_argument("--seed", type=int, default=c.seed)
    p.add_argument("--img_size", type=int, default=c.img_size)
    p.add_argument("--num_classes", type=int, default=c.num_classes)
    p.add_argument("--device", type=str, default="cuda", help="accepted for compatibility; sampling runs on the rank's B200")
    p.add_argument("--slice_size", type=int, default=1, help="accepted for compatibility (training-data option, unused)")
    p.add_argument("--noise_steps", type=int, default=c.noise_steps)
    p.add_argument("--load_model", type=bool, default=True)
    p.add_argument("--img_folder", type=str, default=c.img_folder)
    p.add_argument("--num_samples", type=int, default=c.num_samples, help="number of samples per class")
    p.add_argument("--run_name", type=str, default=c.run_name, help="run name (where to load model)")
    p.add_argument("--dataset_path", type=str, default=c.dataset_path, help="class names = sorted(listdir(<path>/train))")
    p.add_argument("--start_idx", type=int, default=c.start_idx)
    p.add_argument("--sav_denoise_path", type=str, default=c.sav_denoise_path)
    p.add_argument("--vqae_path", type=str, default="models/VQAE/ckpt.pt")
    p.add_argument("--compute_dtype", type=str, default="bf16", choices=["bf16", "f16", "fp32"])
    p.add_argument("--batch_samples", type=int, default=512,
                   help="spectrograms per sampling call (several samp_i are generated together; images do not depend on it)")
    p.add_argument("--colormap", type=str, default="viridis", choices=["viridis", "gray"],
                   help="viridis = matplotlib.cm.viridis as the reference (needs matplotlib); gray = built-in grey ramp "
                        "(same RGBA PNG format; for boxes without matplotlib)")
    p.add_argument("--timing_json", type=str, default=None, help="write this rank's timing summary to <path>.<rank>")
    a = p.parse_args(argv)
    for k, v in vars(a).items():
        setattr(c, k, v)
    return c


def main(argv=None):
    config = parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    set_seed(config.seed)
    if not str(config.device).startswith("cuda"):
        raise SystemExit(f"--device {config.device}: this generator runs on B200 (CUDA) devices only")
    os.makedirs(config.img_folder, exist_ok=True)
    if config.sav_denoise_path is not None:
        os.makedirs(config.sav_denoise_path, exist_ok=True)
    class_names = sorted(os.listdir(os.path.join(config.dataset_path, config.train_folder)))
    diffuser = DiffusionVAE(config.noise_steps, img_size=config.img_size, num_classes=config.num_classes,
                            device=f"cuda:{local}", vqae_path=config.vqae_path, sav_denoise_path=config.sav_denoise_path,
                            class_names=class_names, compute_dtype=config.compute_dtype)
    if config.colormap == "gray":
        import numpy as np

        ramp = np.linspace(0.0, 1.0, 256, dtype=np.float32)
        lut = np.stack([ramp, ramp, ramp, np.ones_like(ramp)], axis=1)
        diffuser.colormap = lambda a: lut[np.asarray(a, dtype=np.uint8)] if np.asarray(a).dtype == np.uint8 \
            else lut[(np.clip(np.asarray(a), 0, 1) * 255).astype(np.uint8)]
    diffuser.prepare(config)
    diffuser.load_model(config)
    # The job is the flattened list of (samp_i, class) pairs: global sample index g = (samp_i - start_idx) * num_classes + i,
    # file {class}_gen_imgs_{i}_{samp_i}.png.  Each rank takes a contiguous slice of it (balanced to one spectrogram) and
    # samples it in batches of --batch_samples; the Philox stream is keyed by (seed, start_idx * num_classes + g), so every
    # image is the same on any rank count and for any batching.
    from .sharding import shard_bounds

    total = config.num_samples * config.num_classes
    lo, hi = shard_bounds(total, world, rank)
    group = max(1, config.batch_samples)
    labels = torch.arange(config.num_classes).long()
    from concurrent.futures import ThreadPoolExecutor

    stats = {"rank": rank, "world": world, "spectrograms": 0, "sample_s": 0.0, "write_s": 0.0, "write_wait_s": 0.0,
             "gpu_launches": 0}

    def timed_write(*a):
        t0 = time.perf_counter()
        paths = diffuser.write_images(*a)
        stats["write_s"] += time.perf_counter() - t0
        return paths

    t_all = time.perf_counter()
    if config.sav_denoise_path:
        # (:765-769) trajectory dumps instead of final images; they are named by class only, so one pass over the classes
        for samp_i in range(config.start_idx + rank, config.start_idx + config.num_samples, world):
            diffuser.gen_images(config.img_folder, samp_i, labels, seed=config.seed,
                                sample_base=samp_i * config.num_classes)
        print("done!")
        return
    with ThreadPoolExecutor(max_workers=1) as writer:  # colour map + PNG encoding of batch k overlap the sampling of k+1
        pending = None
        for first in range(lo, hi, group):
            idx = range(first, min(first + group, hi))
            entries = [(g % config.num_classes, g % config.num_classes, config.start_idx + g // config.num_classes) for g in idx]
            t0 = time.perf_counter()
            images = diffuser.sample(False, torch.tensor([e[0] for e in entries]), seed=config.seed,
                                     sample_base=config.start_idx * config.num_classes + first)
            torch.cuda.synchronize()
            stats["sample_s"] += time.perf_counter() - t0
            stats["gpu_launches"] += diffuser.gpu_launches
            stats["spectrograms"] += len(entries)
            t0 = time.perf_counter()
            if pending is not None:
                pending.result()
            stats["write_wait_s"] += time.perf_counter() - t0
            pending = writer.submit(timed_write, config.img_folder, None, entries, images)
        t0 = time.perf_counter()
        if pending is not None:
            pending.result()
        stats["write_wait_s"] += time.perf_counter() - t0
    stats["total_s"] = time.perf_counter() - t_all
    if config.timing_json:
        with open(f"{config.timing_json}.{rank}", "w") as f:
            json.dump(stats, f)
    print("done!")


if __name__ == "__main__":
    main()
