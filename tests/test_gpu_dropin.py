"""Drop-in check against the REAL reference headers: oracle/_ref/ref_gpu_unit_tests is the reference's own CUDA unit
tests (test/unit_cuda/{cuda/device_vector, primitives/gather, tree/csarray, tree/octree, focus/inject,
domain/domaindecomp_gpu, traversal/groups, traversal/macs, halos/gather_halos_gpu,
primitives/primitives_gpu}.cu, compiled unmodified where they lie under /root/reference) linked against
cornerstone-octree_b200/compat/cstone_gpu_forwarders.cu -> libcstone_b200.so instead of the reference's cstone_gpu
library (recipe: oracle/Makefile; GoogleTest replaced by tests/compat/gtest/gtest.h).  Among others this runs
test/unit_cuda/tree/csarray.cu:164-185 (GPU tree == CPU tree) and test/unit_cuda/tree/octree.cu:25-72 (linked octree
== CPU octree) against this library."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARY = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_unit_tests")


@pytest.mark.skipif(not os.path.exists(BINARY), reason="oracle/_ref/ref_gpu_unit_tests is built where /root/reference exists")
def test_reference_cuda_unit_tests_pass_against_this_library():
    r = subprocess.run([BINARY], capture_output=True, text=True, timeout=300)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    m = re.search(r"(\d+) tests ran, (\d+) failed", out)
    assert m and int(m.group(2)) == 0, out[-2000:]
    ran = set(re.findall(r"\[       OK \] (\S+)", out))
    expected = {"DeviceVector.Construct", "SortByKey.minimal", "CsArrayGpu.computeNodeCountsGpu",
                "CsArrayGpu.rebalanceDecision", "CsArrayGpu.rebalanceTree", "CsArrayGpu.computeOctreeRandom",
                "CsArrayGpu.distributedMockUp", "OctreeGpu.irregularL3", "OctreeGpu.regularL6", "FocusGpu.injectKeysGpu",
                "DomainDecomposition.createSendListGpu", "TargetGroups.t0", "TargetGroups.groupVolumes",
                "Macs.limitSource4x4_matchCPU", "Halos.gatherRanges", "PrimitivesGpu.MinMax"}
    assert expected <= ran, sorted(expected - ran)
