import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only)."""
    from oracle import oracle as orc

    orc.build()
    return orc


@pytest.fixture(scope="session")
def engine():
    """One engine on cuda:0.  No fallback: without the built library or a B200 this errors out."""
    import rbqoc_b200

    eng = rbqoc_b200.Engine(0)
    yield eng
    eng.close()


ENGINE_OPTIONS = ("RBQ_NO_TABLE_CLASS", "RBQ_MC_ZEROCOPY", "RBQ_MC_CHUNK_WAVES", "RBQ_MC_VARIANT", "RBQ_PEAK_MODE",
                  "RBQ_NOISE_GATE_PARALLEL", "RBQ_LINDBLAD_NO_SHARED_MAP", "RBQ_LINDBLAD_TABLE", "RBQ_LINDBLAD_SPT", "RBQ_ROLLOUT_PACKED",
                  "RBQ_ROLLOUT_SCAN", "RBQ_JAC_NT_UNSCENTED", "RBQ_JAC_STRUCTURED", "RBQ_JAC_STAGE_KB", "RBQ_JAC_STAGED",
                  "RBQ_BP_TMA", "RBQ_BP_GENERIC", "RBQ_BP_WARPS")


@pytest.fixture(autouse=True)
def _restore_engine_options(request):
    """Tests flip kernel-selection switches with engine.set_option (the library latches RBQ_* environment variables once,
    in rbq_create); every test leaves the shared engine at its defaults."""
    yield
    if "engine" in request.fixturenames:
        eng = request.getfixturevalue("engine")
        for name in ENGINE_OPTIONS:
            eng.set_option(name, None)
