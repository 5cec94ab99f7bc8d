"""-m gpu, needs 2 GPUs (skipped otherwise): the data-parallel G+D step over NCCL — two ranks, one image each, CUDA kernels,
bucketed all-reduce (ReduceOp.AVG) overlapped with backward — against the §8e target: the mean over ranks of the CPU
oracle's per-shard gradients."""
import os

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_data_parallel_step_matches_mean_of_shard_oracles(tmp_path):
    import test_ddp_gloo as T
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(T._step_worker, args=(2, port, str(tmp_path), "nccl"), nprocs=2, join=True)
    T.check_against_shard_oracles(tmp_path)
