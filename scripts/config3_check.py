"""Config 3 (BASELINE.json): V2 transformer forward + SIP scorer + CRF on synthetic audio, reference vs ours on the same GPU.
Prints, per installation level, how the Note list differs from the unmodified reference's.  usage: python scripts/config3_check.py [seconds]"""
import copy
import sys
import time

import torch

sys.path.insert(0, ".")
from baseline import ref_loader  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0


def key(n):
    return (n.pitch, round(n.start, 4), round(n.end, 4), n.velocity)


def run(model, x, reps=1):
    torch.cuda.synchronize()
    t0 = time.time()
    with torch.no_grad():
        for _ in range(reps):
            notes = model.transcribe(x)
    torch.cuda.synchronize()
    return notes, (time.time() - t0) / reps


def diff(a, b):
    ka, kb = [key(n) for n in a], [key(n) for n in b]
    sa, sb = set(ka), set(kb)
    coarse_a = set((n.pitch, round(n.start, 2), round(n.end, 2)) for n in a)
    coarse_b = set((n.pitch, round(n.start, 2), round(n.end, 2)) for n in b)
    d = dict(n_ref=len(a), n_ours=len(b), identical=ka == kb, only_ref=len(sa - sb), only_ours=len(sb - sa),
             coarse_only_ref=len(coarse_a - coarse_b), coarse_only_ours=len(coarse_b - coarse_a))
    if sa != sb:
        d["examples_ref"] = sorted(sa - sb)[:4]
        d["examples_ours"] = sorted(sb - sa)[:4]
    return d


dev = torch.device("cuda")
model_ref, conf = ref_loader.load_model(dev)
x = torch.from_numpy(ref_loader.synthetic_audio(seconds, seed=seed)).to(dev)
run(model_ref, x)
notes_ref, t_ref = run(model_ref, x)
print(f"reference on {torch.cuda.get_device_name(0)}: {len(notes_ref)} notes, {t_ref*1e3:.0f} ms per {seconds:.0f} s of audio", flush=True)
notes_ref2, _ = run(model_ref, x)
print("reference run-to-run:", diff(notes_ref, notes_ref2), flush=True)

from transkun_b200.transcribe import install_into  # noqa: E402

for label, kw in (("CRF only", dict(patch_scorer=False, patch_frontend=False)),
                  ("CRF + scorer (3xTF32)", dict(patch_scorer=True, patch_frontend=False)),
                  ("CRF + scorer + frontend", dict(patch_scorer=True, patch_frontend=True))):
    m = install_into(copy.deepcopy(model_ref), **kw)
    run(m, x)
    notes, t = run(m, x, reps=3)
    print(f"{label}: {t*1e3:.0f} ms;", diff(notes_ref, notes), flush=True)
torch.backends.cuda.matmul.allow_tf32 = True
m = install_into(copy.deepcopy(model_ref), patch_scorer=True, patch_frontend=True)
run(m, x)
notes, t = run(m, x, reps=3)
print(f"all three, one-pass TF32 scorer (allow_tf32): {t*1e3:.0f} ms;", diff(notes_ref, notes), flush=True)

# ---- batched segment pipeline (transkun_b200.batched): same notes, far fewer launches ----
torch.backends.cuda.matmul.allow_tf32 = False
from transkun_b200.batched import transcribe_batched  # noqa: E402
m = install_into(copy.deepcopy(model_ref), patch_scorer=True, patch_frontend=True)
for mb in (4, 8, 16):
    transcribe_batched(m, x, max_batch=mb)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        notes_b = transcribe_batched(m, x, max_batch=mb)
    torch.cuda.synchronize()
    print(f"batched segments (max_batch={mb}): {(time.time() - t0) / 3 * 1e3:.0f} ms;", diff(notes_ref, notes_b), flush=True)
