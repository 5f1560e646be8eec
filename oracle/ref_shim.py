"""Import shim for the UNMODIFIED reference (test infrastructure only, never the product path).

Usable where /root/reference exists (the build container) or where oracle/vendor_ref.py left its
byte-for-byte copy under baseline/_ref/ (git-ignored; it travels to the GPU box).  Used by
oracle/gen_golden.py to manufacture the fixtures under tests/golden/, by the container-only
cross-checks of the oracle restatements, by bench.py's CPU legs (`--impl reference`, `cpu_baseline`:
oracle/ref_bench.py) and by the one GPU test that runs the reference's own train_ddpg against the
drop-ins.  Nothing in the product package, smoke() or the timed GPU region of bench.py imports it.

What it does:
  * stubs `pymodbus` (reference imports it at WeightEnv.py:7-8, MutiConditionEnv.py:5,
    CommonInterface/modbus_slave.py:4-5; the latter also starts a TCP server thread at import),
  * freezes the wall clock seen by VirtualWeightController (VirtualWeightController.py:1,173-176,230)
    so `elapsed_ms == 1` and the simulator is a pure function of (params, random stream),
  * puts /root/reference and /root/reference/DifferentModules on sys.path without writing bytecode.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
VENDORED_ROOT = os.path.join(os.path.dirname(_HERE), "baseline", "_ref", "erlfill_reference")   # oracle/vendor_ref.py


def _resolve_root():
    """ERLFILL_REFERENCE_ROOT, else /root/reference (build container), else the vendored byte-for-byte copy
    (baseline/_ref/, git-ignored, travels to the GPU box: bench.py's CPU legs and tests/test_ref_train_ddpg_gpu.py)."""
    env = os.environ.get("ERLFILL_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", VENDORED_ROOT):
        if os.path.isfile(os.path.join(cand, "VirtualWeightController.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _resolve_root()


def available() -> bool:
    return os.path.isdir(REFERENCE_ROOT) and os.path.isfile(
        os.path.join(REFERENCE_ROOT, "VirtualWeightController.py"))


def install():
    """Install the stubs and return the frozen-clock VirtualWeightController module."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "pymodbus" not in sys.modules:
        pm, c, d, s = (types.ModuleType(n) for n in
                       ("pymodbus", "pymodbus.client", "pymodbus.datastore", "pymodbus.server"))

        class ModbusSerialClient:
            def __init__(self, *a, **k):
                pass

            def connect(self):
                return False

        class _Ctx:
            def __init__(self, *a, **k):
                pass

        c.ModbusSerialClient = ModbusSerialClient
        d.ModbusSlaveContext = d.ModbusSequentialDataBlock = d.ModbusServerContext = _Ctx
        s.StartTcpServer = lambda *a, **k: None
        sys.modules.update({"pymodbus": pm, "pymodbus.client": c,
                            "pymodbus.datastore": d, "pymodbus.server": s})
    sys.dont_write_bytecode = True
    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "DifferentModules")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import VirtualWeightController as V

    class _FrozenTime:
        time = staticmethod(lambda: 0.0)
        sleep = staticmethod(lambda x: None)

    V.time = _FrozenTime
    return V
