import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_lib():
    """The C-ABI library, loaded; GPU tests fail loudly if it is missing."""
    from hy2dl_b200 import _lib
    return _lib.load()


@pytest.fixture(autouse=True)
def _name_parity_records(request):
    import _util
    _util._current_test[0] = request.node.nodeid
    yield


def pytest_sessionfinish(session, exitstatus):
    """Element-wise parity figures of every comparison made in this session (tests/_util.assert_close)."""
    try:
        import json
        import _util
        if not _util.PARITY_LOG:
            return
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        gpu = any("gpu" in (r["test"] or "") for r in _util.PARITY_LOG)
        worst = sorted(_util.PARITY_LOG, key=lambda r: -r["frac_outside_elementwise"])[:40]
        summary = {"comparisons": len(_util.PARITY_LOG),
                   "with_any_element_outside_allclose": sum(r["frac_outside_elementwise"] > 0 for r in _util.PARITY_LOG),
                   "worst_frac_outside_elementwise": worst[0]["frac_outside_elementwise"],
                   "worst_max_err_over_scale": max(r["max_err_over_scale"] for r in _util.PARITY_LOG),
                   "worst": worst}
        with open(os.path.join(out, "parity_report_gpu.json" if gpu else "parity_report_cpu.json"), "w") as f:
            json.dump(summary, f, indent=1)
        # every comparison, one line each: test | what | n | rtol | max err / scale | share outside allclose | share outside the gate
        with open(os.path.join(out, "parity_all_gpu.tsv" if gpu else "parity_all_cpu.tsv"), "w") as f:
            for r in _util.PARITY_LOG:
                f.write("\t".join([r["test"].split("::")[-1], r["what"], str(r["n"]), f"{r['rtol']:.1e}", f"{r['max_err_over_scale']:.2e}",
                                   f"{r['frac_outside_elementwise']:.2e}", f"{r['frac_outside_gate']:.2e}"]) + "\n")
    except Exception as e:  # a report must never fail the suite
        print("parity report not written:", e)
