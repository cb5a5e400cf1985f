"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares
(no compute calls here)."""

import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "refiners_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    import __graft_entry__ as entry
    from refiners_b200 import backend

    if not backend.is_built():
        entry.build()
    lib = ctypes.CDLL(str(backend.library_path()))
    declared = declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(backend.exported_symbols()) == declared, "ctypes bindings and header disagree"
    assert backend.load_library().rb200_abi_version() == 1


def test_no_product_import_of_oracle():
    """Nothing under refiners_b200/ may import the oracle (it is test infrastructure)."""
    for path in (ROOT / "refiners_b200").rglob("*.py"):
        src = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{path} imports the oracle"
