"""The built library must contain the Blackwell-native instruction forms the design relies on (B200_PROFILING.md,
"What proves a Blackwell-native kernel"): tcgen05 int8 MMAs (UTCIMMA; .2CTA in the CTA-pair energy kernel), TMA loads
(UTMALDG; .2CTA variants whose bytes are accounted on the pair leader's barrier), multicast commits, TMEM loads (LDTM).
CPU-only: reads the SASS of the in-tree .so with cuobjdump, no GPU needed."""
import collections
import re
import shutil
import subprocess

import pytest

from gml_b200 import _lib


def _sass_by_kernel():
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not shutil.which(tool):
        pytest.skip("cuobjdump not available")
    _lib.load()          # builds the library in-tree if it is missing
    out = subprocess.run([tool, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    ops = collections.defaultdict(collections.Counter)
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
        if m and name:
            ops[name][m.group(1)] += 1
    return ops


def test_contraction_kernels_use_tcgen05_tma_tmem():
    ops = _sass_by_kernel()
    pair = [k for k in ops if "tc_energy_pair_kernel" in k]
    grad = [k for k in ops if "tc_grad_kernel" in k]
    stream = [k for k in ops if "tc_energy_kernelI" in k]
    assert len(pair) >= 8 and len(grad) == 6 and len(stream) >= 8    # grad: <1,128> <1,256> <2,128> <2,256> <3,128> <4,128>
    for k in pair:
        c = ops[k]
        assert c["UTCIMMA.2CTA"] > 0, k                       # tcgen05.mma.cta_group::2.kind::i8
        assert c["UTMALDG.2D.2CTA"] > 0, k                    # TMA loads signalling the leader CTA's barrier
        assert c["UTCBAR.2CTA.MULTICAST"] > 0, k              # tcgen05.commit ... multicast::cluster
        assert c["LDTM.x8"] + c["LDTM.x16"] > 0, k            # tcgen05.ld
        assert c["UCGABAR_ARV"] > 0 and c["UCGABAR_WAIT"] > 0, k   # cluster barrier around TMEM alloc / dealloc
    for k in grad + stream:
        c = ops[k]
        assert c["UTCIMMA"] > 0 and c["UTMALDG.2D"] > 0 and c["LDTM.x16"] > 0 and c["UTCBAR"] > 0, k
    # no legacy tensor path anywhere in the library
    for k, c in ops.items():
        assert not any(op.startswith(("HMMA", "IMMA", "HGMMA", "IGMMA")) for op in c), k
