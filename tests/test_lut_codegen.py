"""Codegen guard for the table kernel (CPU-side: reads the SASS of the built library).

The hot loop of ``tdx_lut16_kernel<TabNormal...>`` only runs at ~0.75 of the HBM roof if ptxas keeps
the generator's constants in uniform registers.  It stops doing so -- silently: same source-level
behaviour, same test results, 13 % slower -- when it cannot prove the CTA's warps converged at the
warp reductions in front of the loop (see the comment on ``for_each_listed_chunk`` in
``torchdistx_b200/csrc/kernels/tdx_init_kernels.cu``).  The symptoms are exact: a ``BRA.DIV`` in
front of the ``REDUX`` instructions and local-memory reloads (``LDL``) between the vector stores.
"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "torchdistx_b200", "libtdx_init.so")

# tdx_lut16_kernel<TabNormal<bf16|f16, 10, false>, ..., PKEYS, 7, 7, 4>: the kernels Llama-class models run
HOT = re.compile(r"tdx_lut16_kernelINS_9TabNormalI(13__nv_bfloat16|6__half)Li10ELb0EEES\d_Li10ELb[01]ELi7ELi7ELi4EEE")


def _functions():
    out = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    for chunk in out.split("Function : ")[1:]:
        name, _, body = chunk.partition("\n")
        if HOT.search(name):
            yield name.strip(), [l for l in body.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="needs cuobjdump (CUDA toolkit)")
@pytest.mark.skipif(not os.path.exists(LIB), reason="libtdx_init.so not built")
def test_table_kernel_keeps_its_constants_in_uniform_registers():
    seen = 0
    for name, lines in _functions():
        seen += 1
        assert not any("BRA.DIV" in l for l in lines), f"{name}: ptxas could not prove convergence (BRA.DIV)"
        stores = [i for i, l in enumerate(lines) if "STG.E.EF.128" in l]
        # the fast path of a tile: from its first vector store to the 16th (offset 15 * 1024 * 16 bytes)
        last = [i for i in stores if "+0x3c000]" in lines[i]]
        assert stores and last, f"{name}: hot loop not found"
        hot = lines[stores[0]:last[0] + 1]
        ldl = sum("LDL" in l for l in hot)
        # (the rare k == 0 path, one call site per group of four vectors, may reload a few values; the
        # twin for launches whose descriptors do not share one seed -- PKEYS = false -- keeps the Philox
        # keys in vector registers and has less room)
        pkeys = "Li10ELb1ELi7" in name
        assert ldl <= (8 if pkeys else 16), f"{name}: {ldl} local-memory reloads between the vector stores of a tile"
        # the polynomial's coefficients are uniform-register operands of the FFMAs
        ffma = [l for l in hot if "FFMA" in l]
        assert sum("UR" in l for l in ffma) >= len(ffma) // 2, f"{name}: FFMA constants are not in uniform registers"
    assert seen >= 2, "table kernels not found in libtdx_init.so"
