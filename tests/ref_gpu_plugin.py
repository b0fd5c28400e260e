"""pytest plugin (-p ref_gpu_plugin): the reference's own test modules against the CUDA backend.

Loaded by tests/test_gpu_reference_testsuite.py in a subprocess whose sys.path starts with
baseline/_ref (the unmodified reference + its MDAnalysisTests package).  Before collection it
registers the CUDA backend under the keys the reference's tests parametrise over ('serial',
'openmp'), makes the capped-distance front ends default to it and points FastNS /
CalcRMSDRotationalMatrix at the GPU twins, so every distance, neighbour-search and QCP call of
those test modules is answered by libmdagpu.so.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    if ROOT not in sys.path:
        sys.path.insert(1, ROOT)
    import warnings
    warnings.filterwarnings("ignore")
    import mdanalysis_b200.register as reg
    reg.install(patch_capped=True, masquerade=["serial", "openmp"], capped_default="cuda", patch_classes=True,
                patch_analysis=True)
    from mdanalysis_b200 import _cabi
    config._mdagpu_launches0 = _cabi.launch_count()


def pytest_terminal_summary(terminalreporter):
    from mdanalysis_b200 import _cabi
    n = _cabi.launch_count() - getattr(terminalreporter.config, "_mdagpu_launches0", 0)
    terminalreporter.write_line(f"mdagpu kernel launches during this session: {n}")
