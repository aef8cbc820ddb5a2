"""CPU: the drop-in transcription entry installs transkun_b200.CRF under the reference's module name and delegates to
the reference's own main() with the reference's own flags (checked against a stand-in `transkun` package: the real
one needs moduleconf / pretty_midi / pydub, which are not in this image)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_entry_installs_crf_and_delegates(tmp_path):
    pkg = tmp_path / "transkun"
    pkg.mkdir()
    (pkg / "__init__.py").write_text("")
    (pkg / "ModelTransformer.py").write_text("from . import CRF\n")
    (pkg / "transcribe.py").write_text(textwrap.dedent("""
        import argparse
        def main():
            ap = argparse.ArgumentParser()
            ap.add_argument("audioPath"); ap.add_argument("outPath")
            ap.add_argument("--device", default="cpu", nargs="?")
            ap.add_argument("--segmentHopSize", type=float, required=False)
            args = ap.parse_args()
            from . import ModelTransformer
            print("CRF=" + ModelTransformer.CRF.NeuralSemiCRFInterval.__module__)
            print("ARGS=%s %s %s %s" % (args.audioPath, args.outPath, args.device, args.segmentHopSize))
    """))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(tmp_path), ROOT]))
    out = subprocess.run([sys.executable, "-m", "transkun_b200.transcribe", "in.mp3", "out.mid", "--segmentHopSize", "8"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "CRF=transkun_b200.CRF.NeuralSemiCRFInterval" in out.stdout
    assert "ARGS=in.mp3 out.mid cuda 8.0" in out.stdout   # --device cuda appended: there is no CPU path
    out = subprocess.run([sys.executable, "-m", "transkun_b200.transcribe", "a", "b", "--device", "cuda:1"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert "ARGS=a b cuda:1 None" in out.stdout


def test_entry_reports_missing_reference():
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-m", "transkun_b200.transcribe", "a", "b"], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "delegates to the reference package" in out.stderr


def test_install_rebinds_the_real_reference_modules():
    """With the real reference importable (baseline/_ref), install() makes ModelTransformer construct OUR CRF class,
    OUR scorer class and OUR frontend class; TKB_PATCH_*=0 keeps the reference's torch modules."""
    import pytest
    sys.path.insert(0, ROOT)
    from baseline import ref_loader
    if not ref_loader.available():
        pytest.skip("baseline/_ref (the installed reference) is not present")
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r)
        from baseline import ref_loader
        ref_loader.import_reference()
        from transkun_b200.transcribe import install
        install()
        import transkun.ModelTransformer as MT
        print(MT.CRF.NeuralSemiCRFInterval.__module__, MT.ScaledInnerProductIntervalScorer.__module__, MT.MelSpectrum.__module__)
    """ % ROOT)
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT),
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.split() == ["transkun_b200.CRF.NeuralSemiCRFInterval", "transkun_b200.LayersTransformer",
                                  "transkun_b200.Util"]
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT,
                         env=dict(os.environ, PYTHONPATH=ROOT, TKB_PATCH_SCORER="0", TKB_PATCH_FRONTEND="0"),
                         capture_output=True, text=True, timeout=300)
    assert out.stdout.split() == ["transkun_b200.CRF.NeuralSemiCRFInterval", "transkun.LayersTransformer", "transkun.Util"]


def test_install_into_swaps_modules_and_keeps_the_weights():
    """install_into() on a constructed reference model: the frontend and the scorer become ours and carry the
    checkpoint's tensors bit for bit (same parameter / buffer names)."""
    import pytest
    import torch
    sys.path.insert(0, ROOT)
    from baseline import ref_loader
    if not ref_loader.available():
        pytest.skip("baseline/_ref (the installed reference) is not present")
    model, _ = ref_loader.load_model("cpu")
    before = {k: v.clone() for k, v in model.state_dict().items()}
    from transkun_b200.transcribe import install_into
    install_into(model)
    assert type(model.framewiseFeatureExtractor).__module__ == "transkun_b200.Util"
    assert type(model.scorer).__module__ == "transkun_b200.LayersTransformer"
    after = model.state_dict()
    assert set(after) == set(before)
    for k in before:
        assert torch.equal(before[k], after[k]), k


def test_tf32_split_is_exact():
    """The 3xTF32 operand split: hi has its 13 low mantissa bits clear (TF32-exact) and hi + lo == x exactly."""
    import torch
    from transkun_b200.LayersTransformer import _split_tf32
    x = torch.randn(4096) * torch.logspace(-6, 6, 4096)
    hi, lo = _split_tf32(x)
    assert torch.all((hi.view(torch.int32) & 8191) == 0)
    assert torch.equal(hi + lo, x)
    assert torch.all(lo.abs() <= x.abs() * 2.0 ** -10)
