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


def _build_async(tmp_path):
    exe = str(tmp_path / "spec_rounds_async_model")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "spec_rounds_async_model.cpp")], check=True)
    return exe


def test_spec_rounds_under_arbitrary_interleavings(tmp_path):
    """The stages of the device do not move in lock step: the same protocol with a random scheduler (any stage that has the records it
    needs may go on; stages in front run ahead as far as the flow control of the two exit slots allows; certified stages leave only their
    final record behind) — soundness, no deadlock, no record overwritten before its reader has read it."""
    out = subprocess.run([_build_async(tmp_path), "3000"], capture_output=True, text=True)
    assert out.returncode == 0 and "ok (3000 cases)" in out.stdout, out.stdout[-2000:]


def test_the_async_model_notices_broken_rules(tmp_path):
    exe = _build_async(tmp_path)
    no_flow_control = subprocess.run([exe, "500", "2"], capture_output=True, text=True)
    assert no_flow_control.returncode != 0 and "FAIL" in no_flow_control.stdout
    final_first = subprocess.run([exe, "500", "3"], capture_output=True, text=True)
    assert final_first.returncode != 0 and "FAIL" in final_first.stdout
