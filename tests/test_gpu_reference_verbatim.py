"""The reference's own test suite, run UNMODIFIED against the drop-in package (SURVEY.md section 4,
implication 1): the Monte-Carlo tests and, with the analytic engine on the GPU, the whole of ``test/``.

``oracle/Makefile`` copies ``/root/reference/test`` and ``/root/reference/demo`` into
``oracle/_ref/reference_tests`` (git-ignored, travels to the GPU box, which has no ``/root/reference``).  Here the
files are copied once more into a temporary directory and handed to a fresh ``pytest`` process whose import path
resolves ``mc_dagprop`` to this repository's alias package:

* default generator (Philox contract): every test of ``test_simulator.py``, ``test_monte_carlo_extra.py`` and
  ``test_naming_conventions.py`` except the three known-answer tests that pin the reference's Xoshiro256++ stream
  (``test_simulator.py:139-172``) -- a different stream by design;
* reference-compatible stream (``MCDP_OPT_RNG_STREAM = 1``, switched on by a conftest written next to the copies):
  all of them, the three known-answer tests included;
* the whole reference suite -- all eight files of ``test/`` with ``demo/`` beside them (analytic propagator, PMF
  class, distributions, MC-vs-analytic parity suite, analytic demo) -- under the default generator.
"""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "reference_tests", "test")
MC_FILES = ("test_simulator.py", "test_monte_carlo_extra.py", "test_naming_conventions.py")
XOSHIRO_KAT = (
    "test_simulator.py::TestSimulator::test_empirical_absolute",
    "test_simulator.py::TestSimulator::test_empirical_relative",
    "test_simulator.py::TestSimulator::test_empirical_relative_with_exponential",
)

COMPAT_CONFTEST = '''
# written by tests/test_gpu_reference_verbatim.py: every propagator the tests construct uses the
# reference-compatible generator stream (MCDP_OPT_RNG_STREAM = 4, value 1)
import mc_dagprop
import mc_dagprop.monte_carlo

_Base = mc_dagprop.MonteCarloPropagator


class MonteCarloPropagator(_Base):
    def __init__(self, context, generator):
        super().__init__(context, generator)
        self.set_option(4, 1)


for _m in (mc_dagprop, mc_dagprop.monte_carlo):
    _m.MonteCarloPropagator = MonteCarloPropagator
    _m.Simulator = MonteCarloPropagator
'''

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="oracle/_ref/reference_tests not built")]


def _run(tmp_path, conftest: str, extra_args):
    work = tmp_path / "reference_tests"
    work.mkdir()
    for f in MC_FILES:
        shutil.copy(os.path.join(REF_TESTS, f), work / f)
    (work / "conftest.py").write_text(conftest)
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", *MC_FILES, *extra_args]
    return subprocess.run(cmd, cwd=work, env=env, capture_output=True, text=True, timeout=900)


def _summary(out: str) -> str:
    lines = [ln for ln in out.strip().splitlines() if ln.strip()]
    return lines[-1] if lines else ""


def test_reference_mc_tests_pass_unmodified_philox(tmp_path):
    res = _run(tmp_path, "", [a for k in XOSHIRO_KAT for a in ("--deselect", k)])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "passed" in _summary(res.stdout) and "failed" not in _summary(res.stdout)
    assert "3 deselected" in _summary(res.stdout)


def test_reference_mc_tests_pass_unmodified_reference_stream(tmp_path):
    res = _run(tmp_path, COMPAT_CONFTEST, [])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "passed" in _summary(res.stdout) and "failed" not in _summary(res.stdout)
    assert "deselected" not in _summary(res.stdout)


def test_whole_reference_suite_passes_unmodified(tmp_path):
    """All of the reference's test/ (61 tests) against the drop-in package: Monte-Carlo and analytic engines on the
    GPU, the reference's own conftest (it puts the tree's root on sys.path so that ``import demo`` works)."""
    root = tmp_path / "reference_tree"
    shutil.copytree(os.path.dirname(REF_TESTS), root)  # test/ and demo/
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "test", *[a for k in XOSHIRO_KAT for a in ("--deselect", "test/" + k)]]
    res = subprocess.run(cmd, cwd=root, env=env, capture_output=True, text=True, timeout=1800)
    assert res.returncode == 0, res.stdout[-6000:] + res.stderr[-2000:]
    summary = _summary(res.stdout)
    assert "58 passed" in summary and "3 deselected" in summary and "failed" not in summary, summary
