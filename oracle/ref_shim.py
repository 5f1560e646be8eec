"""Import shim for the UNMODIFIED reference (test infrastructure only, never the product path).

Only usable where /root/reference exists (the build container). It is used by
oracle/gen_golden.py to manufacture the fixtures under tests/golden/ and by the
container-only cross-checks of the oracle restatements. Nothing in bench.py,
smoke() or the -m gpu tests imports this file.

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

REFERENCE_ROOT = os.environ.get("ERLFILL_REFERENCE_ROOT", "/root/reference")


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
