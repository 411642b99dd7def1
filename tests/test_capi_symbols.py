"""The C-ABI library loads on a GPU-less host, exports every symbol include/mww.h declares, and fails
LOUDLY (no CPU fallback) when asked to create a handle without a CUDA device."""

import ctypes
import os
import re

import pytest

from conftest import GOLDEN, ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mww.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mww_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    from microwakeword_b200 import _lib
    L = _lib.lib()
    declared = _header_symbols()
    assert len(declared) >= 14
    for sym in declared:
        assert hasattr(L, sym), "libmww_b200.so does not export %s" % sym
    assert sorted(_lib.EXPORTS) == declared


def test_info_struct_layout_matches_header():
    from microwakeword_b200 import _lib
    # every field of `struct mww_info` in include/mww.h is a 4-byte scalar: count them in the header itself
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "mww.h")).read()
    body = hdr[hdr.index("typedef struct mww_info {"):hdr.index("} mww_info;")]
    fields = re.findall(r"^\s*(?:int32_t|float)\s+(\w+);", body, re.M)
    assert [n for n, _ in _lib.MwwInfo._fields_] == fields and ctypes.sizeof(_lib.MwwInfo) == 4 * len(fields)


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is exercised on GPU-less hosts")
    from microwakeword_b200 import _lib
    from microwakeword_b200.engine import StreamEngine
    with pytest.raises(_lib.MwwError) as e:
        StreamEngine(os.path.join(GOLDEN, "okay_nabu_synth_f32.mww"), n_streams=1)
    assert e.value.code == -3 and "no CUDA device" in str(e.value)
    from microwakeword.inference import Model
    with pytest.raises(_lib.MwwError):
        Model(os.path.join(GOLDEN, "okay_nabu_synth_int8.mww"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "microwakeword_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace("oracle/mixednet.c::", ""), f


def test_host_side_reference_quirks():
    from microwakeword_b200.audio.audio_utils import clip_samples_fed, to_int16
    import numpy as np
    for n in (0, 160, 161, 320, 321, 480, 16000, 16001, 15999):
        idx = chunks = 0
        while idx + 320 < 2 * n:                 # audio_utils.py:56
            idx += 320
            chunks += 1
        assert clip_samples_fed(n) == 160 * chunks
    x = to_int16(np.array([-1.5, -1.0, 0.0, 0.5, 1.0], np.float32))
    assert list(x) == [-32768, -32768, 0, 16384, 32767]      # audio_utils.py:47-48
    with pytest.raises(ValueError):
        to_int16(np.zeros(4, np.int32))
