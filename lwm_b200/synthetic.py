"""Seeded synthetic attention inputs shared by bench.py and the parity tests: the values of (tensor, rank, head) come
from their own generator stream, so any process can regenerate any rank's shard of any head without materialising the
rest — what lets rank r check its result against the oracle at 128K tokens (SURVEY.md §8d synthetic inputs: N(0,1)
generated in fp32, then rounded to bf16)."""
import torch

TENSOR_IDS = {"q": 0, "k": 1, "v": 2, "do": 3}


def head_shard(name, rank, head, rows, D=128, base_seed=1234):
    """[rows, D] float32 holding bf16-representable N(0,1) values of tensor `name`, sequence shard `rank`, head `head`"""
    g = torch.Generator(device="cpu").manual_seed(base_seed + 1000003 * TENSOR_IDS[name] + 10007 * rank + head)
    return torch.randn(rows, D, generator=g, dtype=torch.float32).to(torch.bfloat16).float()


def shard(name, rank, rows, H, D=128, base_seed=1234, dtype=torch.bfloat16):
    """[1, rows, H, D] shard of tensor `name` held by `rank`"""
    out = torch.empty(1, rows, H, D, dtype=dtype)
    for h in range(H):
        out[0, :, h] = head_shard(name, rank, h, rows, D, base_seed).to(dtype)
    return out


def head_global(name, world, head, rows_per_rank, D=128, base_seed=1234):
    """[world*rows_per_rank, D]: one head of the whole (un-sharded) tensor, rank-major = sequence order"""
    return torch.cat([head_shard(name, r, head, rows_per_rank, D, base_seed) for r in range(world)], dim=0)
