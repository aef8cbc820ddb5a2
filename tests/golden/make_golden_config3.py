"""Golden Note list of BASELINE.json config 3: the UNMODIFIED reference (baseline/_ref, shipped checkpoint 2.0.pt) run on
CPU over the synthetic stereo excerpts of baseline/ref_loader.synthetic_audio.  Run in the build container:
    python tests/golden/make_golden_config3.py
writes tests/golden/config3_notes.npz (per excerpt: pitch, start, end, velocity arrays in the reference's output order)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from baseline import ref_loader  # noqa: E402

CASES = {"s16_seed0": (16.0, 0), "s30_seed3": (30.0, 3)}


def main():
    torch.manual_seed(0)
    model, _ = ref_loader.load_model("cpu")
    out = {}
    for name, (seconds, seed) in CASES.items():
        x = torch.from_numpy(ref_loader.synthetic_audio(seconds, seed=seed))
        with torch.no_grad():
            notes = model.transcribe(x)
        out[name + "_pitch"] = np.array([n.pitch for n in notes], dtype=np.int32)
        out[name + "_start"] = np.array([n.start for n in notes], dtype=np.float64)
        out[name + "_end"] = np.array([n.end for n in notes], dtype=np.float64)
        out[name + "_velocity"] = np.array([n.velocity for n in notes], dtype=np.int32)
        print(name, len(notes), "notes")
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "config3_notes.npz"), **out)


if __name__ == "__main__":
    main()
