"""The job the north star names, end to end through the generation driver: `python -m spectrogramgenai_b200.generate` under
torchrun on N GPUs of one box -- NUM_SAMPLES spectrograms per class x 27 classes, all 999 timesteps, VQ / decoder tail and
one RGBA PNG per spectrogram (reference: src/ddpm_conditional_generate.py:105-116, gen_images :759-775).  Checkpoints are
synthetic (random init, saved in the reference's file layout); the colour map is the built-in grey ramp when matplotlib is
absent (same RGBA PNG encoding work).

    python scripts/gen_job.py --gpus 8 --num_samples 37        # 999 spectrograms
Prints one JSON line: wall-clock of the whole torchrun, per-rank sampling / PNG-writer seconds, spectrograms/s.
"""
import argparse
import glob
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--num_samples", type=int, default=37, help="spectrograms per class (x 27 classes)")
    ap.add_argument("--noise_steps", type=int, default=1000)
    ap.add_argument("--work", default="/tmp/sgb200_gen_job")
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    shutil.rmtree(a.work, ignore_errors=True)
    from scripts.synth import write_generation_inputs

    write_generation_inputs(a.work)
    out_dir = os.path.join(a.work, "diffusion_samples")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", "-m", "spectrogramgenai_b200.generate", "--run_name", "DDPM_conditional_VAE",
           "--num_samples", str(a.num_samples), "--noise_steps", str(a.noise_steps), "--img_folder", out_dir,
           "--dataset_path", os.path.join(a.work, "data"), "--vqae_path", os.path.join(a.work, "models", "VQAE", "ckpt.pt"),
           "--colormap", "gray", "--timing_json", os.path.join(a.work, "timing")]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    t0 = time.perf_counter()
    r = subprocess.run(cmd, cwd=a.work, env=env, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        print(r.stdout[-3000:], r.stderr[-3000:])
        raise SystemExit(r.returncode)
    ranks = [json.load(open(p)) for p in sorted(glob.glob(os.path.join(a.work, "timing.*")))]
    pngs = glob.glob(os.path.join(out_dir, "*.png"))
    n = sum(x["spectrograms"] for x in ranks)
    slow = max(ranks, key=lambda x: x["total_s"])
    size = os.path.getsize(pngs[0]) if pngs else 0
    print(json.dumps({
        "what": "generate driver under torchrun: sampling (999 timesteps) + VQ/decoder tail + RGBA PNG per spectrogram",
        "gpus": a.gpus, "classes": 27, "samples_per_class": a.num_samples, "spectrograms": n, "png_files": len(pngs),
        "png_bytes_first": size, "noise_steps": a.noise_steps,
        "torchrun_wall_s": round(wall, 2),
        "job_s_slowest_rank": round(slow["total_s"], 3),
        "spectrograms_per_s_job": round(n / slow["total_s"], 2),
        "spectrograms_per_s_incl_process_startup": round(n / wall, 2),
        "per_rank": [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in x.items()} for x in ranks],
        "png_writer_share_of_slowest_rank": round(slow["write_wait_s"] / slow["total_s"], 4),
        "colormap": "gray ramp (matplotlib absent in this image; the reference uses viridis -- same RGBA encode)"}))
    if not a.keep:
        shutil.rmtree(a.work, ignore_errors=True)


if __name__ == "__main__":
    main()
