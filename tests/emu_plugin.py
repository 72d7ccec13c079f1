"""pytest plugin (-p emu_plugin): run GPU-marked tests against emu_ctx.EmuContext (kernel source on the CPU SIMT emulation)
instead of the CUDA library -- a CPU rehearsal of `pytest -m gpu`.  TEST INFRASTRUCTURE ONLY."""


def pytest_configure(config):
    import femcy_b200.stiffnessMtrx as sm
    from emu_ctx import EmuContext
    sm.Context = EmuContext
    import femcy_b200.conjugateGradientSolver as cgs
    cgs.Context = EmuContext
