"""Golden vectors of the DiffusionVAE decode tail: the UNMODIFIED reference modules (VQEmbeddingEMA, Decoder from
/root/reference/src/diff_modules.py, imported with the matplotlib shim of make_golden.py) run on CPU on seeded
latents, with the synthetic VQAE weights of oracle/vae_oracle.py.  Build container only:
    python tests/golden/make_golden_vae.py      -> tests/golden/golden_vae.npz
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import import_reference  # noqa: E402
from oracle.vae_oracle import HIDDEN, LATENT, N_CODES, make_vqae_state_dict  # noqa: E402

VAE_SEED = 4321
CASES = [("s16", 16, 2, 11), ("s64", 64, 1, 12)]  # tag, latent size, n, input seed


def golden_latents(S, n, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn((n, LATENT, S, S), generator=g) * 0.8  # a few percent of the values leave [-1, 1]


def main():
    torch.set_num_threads(os.cpu_count())
    dm, _ = import_reference()
    sd = make_vqae_state_dict(VAE_SEED)
    codebook = dm.VQEmbeddingEMA(n_embeddings=N_CODES, embedding_dim=LATENT)
    codebook.embedding.copy_(sd["codebook.embedding"])
    decoder = dm.Decoder(input_dim=LATENT, hidden_dim=HIDDEN, output_dim=1)
    decoder.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
    codebook.eval()
    decoder.eval()
    out = {}
    for tag, S, n, seed in CASES:
        x = golden_latents(S, n, seed)
        with torch.inference_mode():
            # exactly the tail of DiffusionVAE.sample (:702-706)
            xc = x.clamp(-1, 1)
            q, _, _, _ = codebook(xc)
            # the indices forward() computes internally (:292-296); encode() (:276-284) assumes another input shape
            idx = torch.argmin(((-torch.cdist(xc.reshape(-1, LATENT), codebook.embedding, p=2)) ** 2).float(), dim=-1)
            y = decoder(q)
            u8 = (((y + 1) / 2) * 255).type(torch.uint8)
        out[f"vae_{tag}_q"] = q.numpy()
        out[f"vae_{tag}_idx"] = idx.reshape(-1).numpy().astype(np.int32)
        out[f"vae_{tag}_y"] = y.numpy()
        out[f"vae_{tag}_u8"] = u8.numpy()
    np.savez_compressed(os.path.join(HERE, "golden_vae.npz"), **out)
    print(json.dumps({k: list(v.shape) for k, v in out.items()}))


if __name__ == "__main__":
    main()
