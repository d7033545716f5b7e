"""CPU checks of the drop-in boundary: libplc.so builds, loads, and exports every symbol include/plc.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "plc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(plc_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    for must in ("plc_cell_fwd", "plc_cell_bwd", "plc_pack_weight", "plc_last_error", "plc_bwd_workspace_bytes"):
        assert must in names


def test_library_builds_and_exports_every_declared_symbol():
    import plconv
    lib_path = plconv.build.build()
    lib = ctypes.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(lib, name), f"libplc.so does not export {name}"
    assert lib.plc_abi_version() == 1


def test_binding_table_matches_header():
    from plconv import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_descriptor_struct_layout_matches_header():
    from plconv import _lib
    assert ctypes.sizeof(_lib.PlcCellDesc) == 8 * 4
    assert [f[0] for f in _lib.PlcCellDesc._fields_] == ["B", "H", "W", "Cin", "Ch", "k", "mode", "has_bias"]


def test_host_side_validation_without_gpu():
    """Descriptor validation is host-only code: it must reject bad problems before touching the device."""
    from plconv import _lib
    lib = _lib.load()
    d = _lib.PlcCellDesc(1, 8, 8, 8, 16, 4, _lib.PLC_MODE_BF16_TC, 1)      # even kernel
    assert lib.plc_packed_weight_bytes(ctypes.byref(d), _lib.PLC_PACK_FWD) == 0
    assert b"odd" in lib.plc_last_error()
    d = _lib.PlcCellDesc(1, 8, 8, 8, 12, 3, _lib.PLC_MODE_BF16_TC, 1)      # Ch % 16
    assert lib.plc_packed_weight_bytes(ctypes.byref(d), _lib.PLC_PACK_FWD) == 0
    d = _lib.PlcCellDesc(1, 8, 8, 64, 64, 3, _lib.PLC_MODE_BF16_TC, 1)
    assert lib.plc_packed_weight_bytes(ctypes.byref(d), _lib.PLC_PACK_FWD) == 256 * 9 * 128 * 2
    assert lib.plc_packed_weight_bytes(ctypes.byref(d), _lib.PLC_PACK_DGRAD) == 128 * 9 * 256 * 2
    d = _lib.PlcCellDesc(1, 8, 8, 5, 7, 3, _lib.PLC_MODE_FP32, 1)          # fp32 mode: any channel count
    assert lib.plc_packed_weight_bytes(ctypes.byref(d), _lib.PLC_PACK_FWD) == 9 * 12 * 28 * 4


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "pl-convlstm-gan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+\.*oracle", txt, flags=re.M), f"{f} imports the oracle"
                assert "import_module(\"oracle" not in txt and "__import__(\"oracle" not in txt
