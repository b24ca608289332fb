"""The documented drop-in switch, exercised against the reference's UNMODIFIED scripts (VERDICT r1 missing #1):
with [alias package, reference checkout] on sys.path, ``scripts/train.py`` and ``scripts/infer.py`` import, the hot-path
names resolve to the sm_100a implementation, everything else (``p2igan_bench.data``, DeepKriging / toy models) falls
through to the reference, ``build_generator_for_inference`` (infer.py:83-106) returns our module and a checkpoint
written from the REFERENCE's own P2IGenerator loads into it.

Needs the reference checkout (present in the build container, absent on the GPU box): skipped otherwise.  Runs in a
subprocess so that the alias package's ``sys.modules`` entries do not leak into the other tests."""
import json
import os
import subprocess
import sys

import pytest

import ref_stubs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import importlib, io, json, os, sys
root, ref = sys.argv[1], sys.argv[2]
sys.path[:0] = [os.path.join(root, "p2i-gan-benchmark_b200"), ref, os.path.join(ref, "scripts"), os.path.join(root, "tests")]
import ref_stubs
ref_stubs.install_stubs()
import torch
train = importlib.import_module("train")          # scripts/train.py, unmodified
infer = importlib.import_module("infer")          # scripts/infer.py, unmodified
import p2igan_bench, p2igan_bench.models as M, p2igan_bench.metrics.metric as MM
out = {}
out["train.build_generator"] = train.build_generator.__module__
out["train.ReconstructionLoss"] = train.ReconstructionLoss.__module__
out["train.gan_loss"] = train.gan_loss.__module__
out["train.RainfallMetricSuite"] = train.RainfallMetricSuite.__module__
out["train.P2IDataModule"] = train.P2IDataModule.__module__
out["infer.P2IGenerator"] = infer.P2IGenerator.__module__
out["infer.DKGenerator"] = infer.DKGenerator.__module__
out["infer.STDKGenerator"] = infer.STDKGenerator.__module__
out["M.SimpleGenerator"] = M.SimpleGenerator.__module__
out["MM.names"] = sorted(n for n in MM.__all__)
cfg = json.load(open(os.path.join(ref, "p2igan_bench/config/p2igan_gan_baseline.json")))
g = infer.build_generator_for_inference(cfg)
out["infer.build_generator_for_inference"] = type(g).__module__
d = train.build_discriminator(cfg)
out["train.build_discriminator"] = type(d).__module__
out["dk"] = type(M.build_generator(json.load(open(os.path.join(ref, "p2igan_bench/config/dk.json"))))).__module__
# a checkpoint in the reference trainer's format, written from the REFERENCE's own modules, loads into ours
refM = importlib.import_module("p2igan_bench.models._reference")
torch.manual_seed(7)
g_ref, d_ref = refM.P2IGenerator(cfg), refM.P2IDiscriminator(in_channels=16)
buf = io.BytesIO()
torch.save({"epoch": 1, "global_step": 5, "generator": g_ref.state_dict(), "discriminator": d_ref.state_dict()}, buf)
buf.seek(0)
ck = torch.load(buf, map_location="cpu", weights_only=True)      # infer.py:183-185
state = ck.get("generator", ck)
res = g.load_state_dict(state)
out["load.missing"], out["load.unexpected"] = list(res.missing_keys), list(res.unexpected_keys)
out["load.equal"] = all(torch.equal(v, g.state_dict()[k]) for k, v in state.items())
res = d.load_state_dict(ck["discriminator"])
out["loadD.missing"], out["loadD.unexpected"] = list(res.missing_keys), list(res.unexpected_keys)
# and the other way round
res = g_ref.load_state_dict(g.state_dict())
out["back.missing"], out["back.unexpected"] = list(res.missing_keys), list(res.unexpected_keys)
try:
    M.NoSuchModel
    out["attr_error"] = False
except AttributeError:
    out["attr_error"] = True
print("RESULT " + json.dumps(out))
"""


@pytest.mark.skipif(not ref_stubs.reference_available(), reason="reference checkout not present (GPU box)")
def test_reference_scripts_import_with_alias_first():
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT, ref_stubs.REFERENCE_ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    out = json.loads(line[7:])
    # hot path -> sm_100a implementation
    assert out["train.build_generator"] == "p2igan_bench.models"
    assert out["infer.build_generator_for_inference"] == "p2igan_b200.generator"
    assert out["train.build_discriminator"] == "p2igan_b200.discriminator"
    assert out["infer.P2IGenerator"] == "p2igan_b200.generator"
    assert out["train.ReconstructionLoss"] == "p2igan_b200.losses" and out["train.gan_loss"] == "p2igan_b200.losses"
    assert out["train.RainfallMetricSuite"] == "p2igan_b200.metrics"
    # everything else -> the reference's own files
    assert out["train.P2IDataModule"] == "p2igan_bench.data.dataloader"
    assert out["infer.DKGenerator"].endswith("_reference.dk") and out["infer.STDKGenerator"].endswith("_reference.stdk")
    assert out["M.SimpleGenerator"].endswith("_reference.simple") and out["dk"].endswith("_reference.dk")
    assert out["MM.names"] == sorted(["RegressionMetrics", "CategoricalMetrics", "FractionalSkillScoreMetric", "MetricConfig",
                                      "RainfallMetricSuite"])
    # reference-format checkpoints load both ways with identical keys
    for k in ("load.missing", "load.unexpected", "loadD.missing", "loadD.unexpected", "back.missing", "back.unexpected"):
        assert out[k] == [], (k, out[k])
    assert out["load.equal"] and out["attr_error"]


def test_alias_package_without_reference_still_serves_the_hot_path():
    """No reference on the path (the GPU box): the hot-path names import; names outside it raise a clear AttributeError."""
    code = ("import sys, os; sys.path.insert(0, os.path.join(sys.argv[1], 'p2i-gan-benchmark_b200'));"
            "os.chdir('/'); import p2igan_bench.models as M, p2igan_bench.modules as MD, p2igan_bench.metrics as MT;"
            "assert M.P2IGenerator.__module__ == 'p2igan_b200.generator' and MD.gan_loss and MT.RainfallMetricSuite;\n"
            "try:\n    M.DKGenerator\n    raise SystemExit(3)\nexcept AttributeError as e:\n    assert 'reference checkout' in str(e)")
    env = dict(os.environ)
    env.pop("P2I_REFERENCE_ROOT", None)
    env["PYTHONPATH"] = ""
    r = subprocess.run([sys.executable, "-c", code, ROOT], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
