"""Golden vectors at the 256x256-spectrogram geometry (BASELINE configs[4]; Diffusion's default img_size=256, c_in=1,
/root/reference/src/diff_modules.py:376-378), which the small fixtures of make_golden.py do not reach.

    python tests/golden/make_golden_large.py        (build container only; ~5 min, ~20 GB of host memory)

  * p128c1: the UNMODIFIED reference UNet_conditional(1, 1, num_classes=27) at S = 128, n = 1 (sa1/sa5 L = 4096,
    sa6 L = 16384: nn.MultiheadAttention materialises 4 x 16384^2 fp32 weights = 4.3 GB, which still fits here).
  * p256c1: S = 256, n = 1 (sa6 L = 65536: the reference would need 68 GB for the attention weights alone), computed by
    the oracle restatement with query-chunked attention (oracle/ddpm_oracle.py ATTN_QUERY_CHUNK); the same oracle is
    pinned to the reference at p128c1 / p32c1 / r16 / r64 by tests/test_oracle_golden.py.
Only eps is stored (fp32 [1,1,S,S]); weights and inputs are rebuilt from seeds.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import NUM_CLASSES, WEIGHT_SEED, golden_inputs, import_reference  # noqa: E402
from oracle import ddpm_oracle as O  # noqa: E402
from oracle.weights import make_state_dict  # noqa: E402

LARGE_CASES = [("p128c1", 128, 1, 1, 400, "reference"), ("p256c1", 256, 1, 1, 400, "oracle")]


def main():
    torch.set_num_threads(os.cpu_count())
    dm, _ = import_reference()
    out, info = {}, {}
    for tag, s, c, n, tv, how in LARGE_CASES:
        sd = make_state_dict(WEIGHT_SEED, c, c, NUM_CLASSES)
        x, y = golden_inputs(s, c, n)
        t = (torch.ones(n) * tv).long()
        t0 = time.time()
        if how == "reference":
            m = dm.UNet_conditional(c, c, num_classes=NUM_CLASSES)
            m.load_state_dict(sd, strict=True)
            m.eval()
            with torch.inference_mode():
                out[f"eps_{tag}_t{tv}_cond"] = m(x, t, y).numpy()
                out[f"eps_{tag}_t{tv}_uncond"] = m(x, t, None).numpy()
            ora = O.unet_forward(sd, x, t, y).numpy()
            info[f"{tag}_oracle_vs_reference_rel_l2"] = O.rel_l2(torch.from_numpy(ora), torch.from_numpy(out[f"eps_{tag}_t{tv}_cond"]))
        else:
            out[f"eps_{tag}_t{tv}_cond"] = O.unet_forward(sd, x, t, y).numpy()
        info[f"{tag}_seconds"] = round(time.time() - t0, 1)
        print(tag, info, flush=True)
    np.savez_compressed(os.path.join(HERE, "golden_large.npz"), **out)
    with open(os.path.join(HERE, "golden_large_info.json"), "w") as f:
        json.dump(info, f, indent=1)
    print(json.dumps({"n_arrays": len(out), "bytes": sum(int(v.nbytes) for v in out.values())}))


if __name__ == "__main__":
    main()
