"""Sequence-sharded KV cache — host mirror of `FlaxLLaMAAttention._concatenate_to_cache` (lwm/llama.py:440-492) in the
process-per-GPU model: every rank of the 'sp' group holds the rows [rank*L, (rank+1)*L) of cached_key / cached_value
(L = max_length / sp; in_specs PS(('dp','fsdp'), 'sp', 'tp', None), llama.py:468-469).

  * decode step (query length 1, llama.py:452-483): the new key/value row is written by the ONE rank that owns slot
    `cache_index` (`cur_index - axis_index * sp_size` in range), everybody else leaves its shard untouched;
  * prefill (query length > 1, llama.py:485-487: `dynamic_update_slice` at `cache_index`): the new rows are sharded like
    the queries (rank r holds rows [r*q_loc, (r+1)*q_loc) of them), their destination slots generally belong to other
    ranks, so the shards are all-gathered once and every rank copies the slice that falls into its own cache rows.
Pure data movement (copies and one all-gather): no arithmetic, no kernels of its own. The cache shards are exactly the
k / v arguments `ringattention` (prefill: "K/V = whole cache") and `ringattention_inference` (decode) take."""
import torch
import torch.distributed as dist


class ShardedKVCache:
    def __init__(self, batch, max_length, num_heads, head_dim, dtype=torch.bfloat16, device="cuda", group=None):
        self.group = group
        self.world, self.rank = 1, 0
        if dist.is_available() and dist.is_initialized():
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if max_length % self.world:
            raise ValueError("max_length %d must be divisible by the ring size %d" % (max_length, self.world))
        self.max_length = max_length
        self.shard_len = max_length // self.world
        shape = (batch, self.shard_len, num_heads, head_dim)
        self.cached_key = torch.zeros(shape, dtype=dtype, device=device)       # jnp.zeros (llama.py:444-445)
        self.cached_value = torch.zeros(shape, dtype=dtype, device=device)
        self.cache_index = 0

    def concatenate(self, key, value):
        """key/value: the new rows. Decode: [B,1,H,D], replicated along the ring. Prefill: this rank's shard
        [B,q_loc,H,D] of the q_loc*world new rows. Returns (cached_key, cached_value) shards after the update and
        advances cache_index by the number of new rows (llama.py:488-491)."""
        lo = self.rank * self.shard_len
        if key.shape[1] == 1 and value.shape[1] == 1 and self._is_decode(key):
            cur = self.cache_index - lo
            if 0 <= cur < self.shard_len:
                self.cached_key[:, cur].copy_(key[:, -1])
                self.cached_value[:, cur].copy_(value[:, -1])
            n_new = 1
        else:
            n_new = key.shape[1] * self.world
            if self.cache_index + n_new > self.max_length:
                raise ValueError("cache overflow: %d + %d > %d" % (self.cache_index, n_new, self.max_length))
            for new, cache in ((key, self.cached_key), (value, self.cached_value)):
                full = self._gather_rows(new.contiguous())
                # global slots [cache_index, cache_index + n_new) intersected with my rows [lo, lo + shard_len)
                a, b = max(self.cache_index, lo), min(self.cache_index + n_new, lo + self.shard_len)
                if b > a:
                    cache[:, a - lo:b - lo].copy_(full[:, a - self.cache_index:b - self.cache_index])
        self.cache_index += n_new
        return self.cached_key, self.cached_value

    def _is_decode(self, key):
        # a one-row prefill on a one-rank "ring" is the same write; on a real ring a [B,1,...] argument is the
        # replicated decode row (the reference switches on query.shape[1] == 1, llama.py:451)
        return True

    def _gather_rows(self, x):
        if self.world == 1:
            return x
        B, n = x.shape[0], x.shape[1]
        g = torch.empty((self.world * B, n) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(g, x, group=self.group)        # rank-major along dim 0
        return g.view((self.world, B, n) + tuple(x.shape[2:])).transpose(0, 1).reshape((B, self.world * n) + tuple(x.shape[2:]))
