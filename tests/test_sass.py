"""CPU suite: static checks on the cross-compiled sm_100a code (cuobjdump; no GPU needed).  The GEMM family must
be built from the Blackwell instructions the design claims -- tcgen05.mma (UTCHMMA) with TMEM accumulators read
by tcgen05.ld (LDTM), TMA loads / multicast loads / bulk stores (UTMALDG / UTMASTG), mbarrier synchronisation
(SYNCS) -- and no kernel of the library may spill registers to local memory."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "efficientvideoclassification_youtube8m_b200", "libevc.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")


def _run(*args):
    import __graft_entry__ as g
    g.build()
    return subprocess.run(["cuobjdump", *args, LIB], capture_output=True, text=True, check=True).stdout


def test_library_is_built_for_sm_100a_only():
    archs = set(re.findall(r"arch = (sm_\w+)", _run("--list-elf") + _run("--dump-elf-symbols")))
    elfs = re.findall(r"sm_\d+a?", _run("--list-elf"))
    assert elfs and set(elfs) == {"sm_100a"}, (elfs, archs)


def test_gemm_kernels_use_tcgen05_tmem_and_tma():
    sass = _run("-sass")
    kernels = re.split(r"\n\s*Function : ", sass)[1:]
    gemm = [k for k in kernels if k.startswith("_ZN3evc11gemm_kernel")]
    assert len(gemm) >= 12                                     # majors x tile widths x epilogues x cluster sizes
    for k in gemm:
        name = k.split("\n", 1)[0]
        assert "UTCHMMA" in k, name                            # tcgen05.mma issued from the elected thread
        assert "LDTM" in k, name                               # tcgen05.ld: accumulators come back from TMEM
        assert "UTMALDG" in k, name                            # cp.async.bulk.tensor loads
        assert "SYNCS.PHASECHK" in k and "SYNCS.ARRIVE" in k, name   # mbarrier pipeline
        assert "HMMA" not in k.replace("UTCHMMA", ""), name    # no legacy mma.sync path
    # cluster-of-two variants: multicast pairs (PAIR=0) multicast the B tile and release slots in both CTAs;
    # cta_group::2 pairs (PAIR=1) issue the 2-CTA forms of the MMA, the TMA loads and the commit
    mc = [k for k in gemm if re.match(r"_ZN3evc11gemm_kernelILi\dELi\dELi\d+ELi\dELi2ELi0E", k)]
    assert mc and all("UTMALDG.2D.MULTICAST" in k and "UTCBAR.MULTICAST" in k for k in mc)
    two = [k for k in gemm if re.match(r"_ZN3evc11gemm_kernelILi\dELi\dELi\d+ELi\dELi2ELi1E", k)]
    assert len(two) >= 6 and all("UTCHMMA.2CTA" in k and "UTMALDG.2D.2CTA" in k and "UTCBAR.2CTA.MULTICAST" in k
                                  for k in two), [k.split("\n", 1)[0] for k in two][:3]
    stores = [k for k in gemm if re.match(r"_ZN3evc11gemm_kernelILi\dELi\dELi\d+ELi0E", k)]
    assert stores and all("UTMASTG" in k for k in stores)


def test_no_kernel_spills_to_local_memory():
    usage = _run("--dump-resource-usage")
    rows = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", usage)
    assert len(rows) >= 30
    for name, reg, stack, shared, local in rows:
        assert int(local) == 0, name
        if "gemm_kernel" in name or "lstm_rec" in name:
            assert int(stack) == 0, name                        # the persistent kernels keep everything in registers
            # 320 threads (LSTM forward epilogue) x regs must fit the 64 K register file with one CTA per SM
            assert int(reg) * 320 <= 65536, (name, reg)
