"""GPU tier: BASELINE.json configs[0] -- learned clauses + trail snapshots recorded from the
reference's CPU solver, replayed through the C ABI; per-run hit triples must equal the committed
golden hits (and the reference's own GPU checker must agree when oracle/_ref is present)."""
import pytest

import ref_lib
from golden_replay import as_lists, lib_run, load, replay
from gpusharesat_b200 import GpuClauseSharer, GpuClauseSharerOptions

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("grid", [(3, 32), (-1, -1)])
def test_cuda_path_reproduces_config1_golden_hits(grid):
    fx = load()
    sh = GpuClauseSharer(GpuClauseSharerOptions(gpuBlockCountGuideline=grid[0], gpuThreadsPerBlockGuideline=grid[1],
                                                minGpuLatencyMicros=0, initReportCountPerCategory=4))
    sh.setVarCount(fx["nvars"])
    sh.setCpuSolverCount(fx["nsolvers"])
    runs = replay(fx["events"], sh, lib_run)
    assert as_lists(runs) == fx["expected_hits_per_run"]


@pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built")
def test_reference_gpu_checker_reproduces_config1_golden_hits():
    fx = load()
    sh = ref_lib.RefSharer(blocks=3, threads=32, report=5000)
    sh.setVarCount(fx["nvars"])
    sh.setCpuSolverCount(fx["nsolvers"])
    runs = replay(fx["events"], sh, lib_run)
    assert as_lists(runs) == fx["expected_hits_per_run"]
