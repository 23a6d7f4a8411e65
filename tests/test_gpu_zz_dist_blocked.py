"""The blocked variant of the sharded factorisation (GPK_DIST_OZAKI=1: panels collected in blocks of 8, one sliced int8
update of the block-cyclic columns per block; active from 32 panels, i.e. the N=4096 and the ragged N=4500 cases of the
worker) against the single-GPU evaluation.  Runs last: the variant is new and not the default yet."""
import pytest

pytestmark = pytest.mark.gpu

from test_gpu_dist import _check_sharded  # noqa: E402


@pytest.mark.parametrize("world", [1, 2])
def test_blocked_sharded_eval_matches_single_gpu(world, tmp_path, golden):
    _check_sharded(world, 1, tmp_path, golden)
