"""profiles/rNN_grad_parity.md from the JSON reports tests/test_gpu_train.py writes when DGP_GRAD_REPORT=<dir>/grad_report.json."""
import json
import os
import sys

import numpy as np


def main(prefix, out_path):
    out = ["# Gradient parity of the training step vs torch autograd through the fp32 oracle (tests/test_gpu_train.py)", "",
           "3 frames of 64x96, nj = 4, gm2 = 1, gm3 = 3, chain skeleton; every trainable variable (163 tensors).", "",
           "| storage | leaf | tensors | cosine min | cosine median | rel-L2 median | rel-L2 max |", "|---|---|---|---|---|---|---|"]
    for tag, label in (("fp32_oracle", "bf16"), ("fp16_mode", "fp16")):
        p = "%s_%s.json" % (prefix, tag)
        if not os.path.exists(p):
            continue
        by = {}
        for x in json.load(open(p)):
            by.setdefault(x["name"].split("/")[-1], []).append(x)
        for leaf, v in by.items():
            out.append("| %s | %s | %d | %.5f | %.5f | %.4f | %.4f |" % (
                label, leaf, len(v), min(a["cos"] for a in v), np.median([a["cos"] for a in v]),
                np.median([a["rel_l2"] for a in v]), max(a["rel_l2"] for a in v)))
    out += ["", "bf16 per-variable detail (listed in forward order; the loss sits after the heads):", "",
            "| variable | cosine | rel-L2 | ref norm |", "|---|---|---|---|"]
    for x in json.load(open(prefix + "_fp32_oracle.json")):
        if x["name"].endswith("weights"):
            out.append("| %s | %.5f | %.4f | %.3g |" % (x["name"], x["cos"], x["rel_l2"], x["ref_norm"]))
    out += ["", "Forward tracking of a bf16-rounding emulation of the oracle (tools/diag_emulation.py): GPU vs emulation rms 1e-5 of max",
            "in block 1, growing to 8.7e-4 in block 4 (GPU vs fp32 oracle: 9.9e-4) -- rounding-boundary flips decorrelate the two forwards,",
            "so the emulation cannot serve as a tighter reference; its own gradients differ from the fp32 oracle's by median 8.3 % / max 15.1 %,",
            "the same figures the CUDA path shows.  fp16 storage (8x finer mantissa) brings the CUDA path to median 2.9 % / max 5.3 %."]
    open(out_path, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
