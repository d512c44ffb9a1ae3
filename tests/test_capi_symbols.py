"""The C-ABI shared library builds, loads on a CPU-only box and exports every symbol include/smrt_dort_b200.h declares
(no compute call is made here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from smrt_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_library():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "smrt_b200", "csrc")], check=True)
    return capi.library_path()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "smrt_dort_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smrtb200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(capi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(built_library):
    lib = C.CDLL(built_library)
    for sym in declared_symbols():
        assert hasattr(lib, sym), f"{sym} is declared in the header but not exported"
    assert lib.smrtb200_abi_version() == capi.ABI_VERSION


def test_library_is_sm100a_only(built_library):
    out = subprocess.run(["cuobjdump", "-lelf", built_library], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layout_matches_header():
    """ctypes mirrors of smrtb200_options / smrtb200_batch have the C layout (checked against a compiled probe)."""
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "smrt_dort_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(smrtb200_options), offsetof(smrtb200_options, prune_deep_snowpack),
             offsetof(smrtb200_options, chunk), sizeof(smrtb200_batch), offsetof(smrtb200_batch, phi),
             offsetof(smrtb200_batch, values), offsetof(smrtb200_batch, status));
      return 0;
    }'''
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "probe.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(tmp, "probe"),
                        os.path.join(tmp, "probe.c")], check=True)
        vals = [int(x) for x in subprocess.run([os.path.join(tmp, "probe")], capture_output=True,
                                               text=True).stdout.split()]
    assert vals == [C.sizeof(capi.Options), capi.Options.prune_deep_snowpack.offset, capi.Options.chunk.offset,
                    C.sizeof(capi.Batch), capi.Batch.phi.offset, capi.Batch.values.offset, capi.Batch.status.offset]


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(capi, "_LIB", None)
    monkeypatch.setattr(capi, "library_path", lambda: str(tmp_path / "nope.so"))
    from smrt_b200.error import SMRTError

    with pytest.raises(SMRTError, match="no CPU fallback"):
        capi.load_library()
