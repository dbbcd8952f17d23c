"""CPU model of the speculative rounds (tests/spec_rounds_model.cpp): prediction, correction, certification and the bounded simulations
re-derived from DESIGN.md 4.5 and checked on thousands of random inventories / request mixes / tables against the sequential recurrence —
soundness of the certification, progress, termination within stages + 2 rounds, the correction leaving a consistent prefix alone, no
certification of a cut-off log.  (The device code is checked against the oracle by the -m gpu tests; this checks the PROTOCOL.)"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_spec_rounds_protocol_model(tmp_path):
    exe = str(tmp_path / "spec_rounds_model")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "spec_rounds_model.cpp")], check=True)
    out = subprocess.run([exe, "2500"], capture_output=True, text=True)
    assert out.returncode == 0 and "ok (2500 cases)" in out.stdout, out.stdout[-2000:]


def test_the_frontier_must_be_exempt_from_the_cut_off(tmp_path):
    """With the exemption tied to the lagging knowledge alone ("every stage in front was consistent") the frontier itself gets cut off
    and the rounds exceed stages + 2 — the model shows it; the kernel uses the rule that passes."""
    exe = str(tmp_path / "spec_rounds_model")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "spec_rounds_model.cpp")], check=True)
    out = subprocess.run([exe, "200"], capture_output=True, text=True, env=dict(os.environ, SPEC_MODEL_STRICT_KNOWN="1"))
    assert out.returncode != 0 and "no termination" in out.stdout
