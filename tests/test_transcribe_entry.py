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
    """With the real reference importable (this container only), install() makes ModelTransformer construct OUR
    CRF class and, on request, OUR scorer class."""
    import pytest
    if not os.path.isdir("/root/reference/transkun"):
        pytest.skip("reference checkout not present")
    code = textwrap.dedent("""
        import sys, types
        sys.path.insert(0, "/root/reference")
        for m in ("pretty_midi", "mir_eval", "mir_eval.transcription", "mir_eval.transcription_velocity"):
            sys.modules[m] = types.ModuleType(m)   # not installed here; unused on this path (SURVEY.md section 8c)
        from transkun_b200.transcribe import install
        install(patch_scorer=True)
        import transkun.ModelTransformer as MT
        print(MT.CRF.NeuralSemiCRFInterval.__module__, MT.ScaledInnerProductIntervalScorer.__module__)
    """)
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT),
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.split() == ["transkun_b200.CRF.NeuralSemiCRFInterval", "transkun_b200.LayersTransformer"]
