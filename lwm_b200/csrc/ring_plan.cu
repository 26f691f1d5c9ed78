// The schedule and the heap layout of the peer-memory ring (what replaces the reference's lax.scan + lax.ppermute loop of
// the un-vendored `ringattention` package, entered at lwm/llama.py:539-569) as PURE FUNCTIONS behind the C ABI, so that a
// host in any language can drive the lwm_ring_* primitives: lwm_ring_plan says which query chunks this rank computes,
// which K/V chunks it pulls in which order and group, which tile launches follow each group and which dK/dV partials
// will land in its heap; lwm_ring_layout says where everything lives inside every rank's heap. Host-only code (no
// device is touched). lwm_b200/ring_schedule.py::make_peer_plan and ring_peer.Layout are the same functions in Python;
// tests/test_ring_plan_native_cpu.py compares the two field by field.
#include <string.h>
#include "capi_internal.h"
#include "../../include/lwm_b200.h"

namespace {

struct Ch { int owner, index; long long start, length, pos0; };

bool visible(long long q_pos0, long long q_len, long long k_pos0, int causal) {
  return !causal || (q_pos0 + q_len - 1 >= k_pos0);
}

// query chunks computed by `rank` (references into the contiguous shards)
int compute_chunks(int world, int rank, long long Sq, int zigzag, Ch* out) {
  if (!zigzag) {
    out[0] = Ch{rank, 0, 0, Sq, (long long)rank * Sq};
    return 1;
  }
  const long long h = Sq / 2;
  const int cs[2] = {rank, 2 * world - 1 - rank};
  for (int i = 0; i < 2; ++i) {
    const int c = cs[i];
    out[i] = Ch{c / 2, c % 2, (c % 2) * h, h, (long long)c * h};
  }
  return 2;
}

int kv_chunks_of(int rank, long long Sk, int zigzag, Ch* out) {
  if (!zigzag) {
    out[0] = Ch{rank, 0, 0, Sk, (long long)rank * Sk};
    return 1;
  }
  const long long h = Sk / 2;
  out[0] = Ch{rank, 0, 0, h, (long long)rank * Sk};
  out[1] = Ch{rank, 1, h, h, (long long)rank * Sk + h};
  return 2;
}

// chunks (of all owners, owner-major) that any query chunk of rank r can see
int needed_by(int world, int r, long long Sq, long long Sk, int causal, int zigzag, Ch* out) {
  Ch qs[2];
  const int nq = compute_chunks(world, r, Sq, zigzag, qs);
  int n = 0;
  for (int o = 0; o < world; ++o) {
    Ch cs[2];
    const int nc = kv_chunks_of(o, Sk, zigzag, cs);
    for (int i = 0; i < nc; ++i) {
      bool any = false;
      for (int q = 0; q < nq; ++q) any = any || visible(qs[q].pos0, qs[q].length, cs[i].pos0, causal);
      if (any) out[n++] = cs[i];
    }
  }
  return n;
}

lwm_ring_chunk to_abi(const Ch& c) {
  lwm_ring_chunk o;
  o.owner = c.owner; o.index = c.index; o.start = c.start; o.length = c.length; o.pos0 = c.pos0;
  return o;
}

// one launch per (query chunk, K/V owner): the visible chunks of one owner are adjacent in position and merge
int launches_for(const Ch* chunks, int n, const Ch* qs, int nq, int causal, lwm_ring_launch* out) {
  int owners[LWM_RING_MAX_CHUNKS], n_own = 0;
  for (int i = 0; i < n; ++i) {
    bool seen = false;
    for (int j = 0; j < n_own; ++j) seen = seen || owners[j] == chunks[i].owner;
    if (!seen) owners[n_own++] = chunks[i].owner;
  }
  int m = 0;
  for (int q = 0; q < nq; ++q)
    for (int oi = 0; oi < n_own; ++oi) {
      // the owner's visible chunks sorted by position, merged into contiguous ranges
      Ch v[2];
      int nv = 0;
      for (int i = 0; i < n; ++i)
        if (chunks[i].owner == owners[oi] && visible(qs[q].pos0, qs[q].length, chunks[i].pos0, causal)) v[nv++] = chunks[i];
      if (nv == 2 && v[1].pos0 < v[0].pos0) { Ch t = v[0]; v[0] = v[1]; v[1] = t; }
      int i = 0;
      while (i < nv) {
        long long p0 = v[i].pos0, rows = v[i].length;
        while (i + 1 < nv && v[i + 1].pos0 == p0 + rows) { rows += v[i + 1].length; ++i; }
        out[m].q_chunk = q; out[m].owner = owners[oi]; out[m].key_row0 = p0; out[m].rows = rows;
        ++m;
        ++i;
      }
    }
  return m;
}

long long al256(long long n) { return (n + 255) / 256 * 256; }

}  // namespace

extern "C" int lwm_ring_plan(int world, int rank, long long Sq, long long Sk, int causal, int zigzag, int fwd_group_chunks,
                             lwm_ring_plan_t* out) {
  if (!out) return lwm_fail(LWM_ERR_ARG, "ring_plan: null out");
  if (world < 1 || world > LWM_RING_MAX_WORLD || rank < 0 || rank >= world || Sq <= 0 || Sk <= 0 || fwd_group_chunks < 1)
    return lwm_fail(LWM_ERR_ARG, "ring_plan: bad arguments");
  if (zigzag && !(Sq == Sk && Sq % 256 == 0))
    return lwm_fail(LWM_ERR_SHAPE, "ring_plan: zigzag layout needs Sq == Sk and a shard length divisible by 256");
  memset(out, 0, sizeof(*out));
  out->world = world; out->rank = rank; out->zigzag = zigzag ? 1 : 0; out->chunks_per_rank = zigzag ? 2 : 1;
  Ch qs[2];
  const int nq = compute_chunks(world, rank, Sq, zigzag, qs);
  out->n_q = nq;
  for (int i = 0; i < nq; ++i) out->q[i] = to_abi(qs[i]);
  for (int peer = 0; peer < world; ++peer) {
    if (peer == rank) continue;
    Ch pq[2];
    const int np = compute_chunks(world, peer, Sq, zigzag, pq);
    for (int i = 0; i < np; ++i)
      if (pq[i].owner == rank) {
        out->q_sends[out->n_q_sends].start = pq[i].start;
        out->q_sends[out->n_q_sends].length = pq[i].length;
        out->q_sends[out->n_q_sends].peer = peer;
        ++out->n_q_sends;
      }
  }
  Ch need[LWM_RING_MAX_CHUNKS], local[2], remote[LWM_RING_MAX_CHUNKS];
  const int n_need = needed_by(world, rank, Sq, Sk, causal, zigzag, need);
  int n_local = 0, n_remote = 0;
  for (int i = 0; i < n_need; ++i)
    if (need[i].owner == rank) local[n_local++] = need[i];
  for (int d = 1; d < world; ++d) {   // ring order: owners rank-1, rank-2, ...
    const int o = ((rank - d) % world + world) % world;
    for (int i = 0; i < n_need; ++i)
      if (need[i].owner == o) remote[n_remote++] = need[i];
  }
  // forward groups: local chunks first, then the remote chunks in groups of fwd_group_chunks (a first remote group of
  // one chunk when nothing is local)
  auto add_group = [&](const Ch* cs, int n, bool fwd) {
    int& ng = fwd ? out->n_fwd_groups : out->n_bwd_groups;
    int* first_chunk = fwd ? out->fwd_group_first_chunk : out->bwd_group_first_chunk;
    int* first_launch = fwd ? out->fwd_group_first_launch : out->bwd_group_first_launch;
    lwm_ring_chunk* chunks = fwd ? out->fwd_chunks : out->bwd_chunks;
    lwm_ring_launch* launches = fwd ? out->fwd_launches : out->bwd_launches;
    int& n_chunks = fwd ? out->n_fwd_chunks : out->n_bwd_chunks;
    int& n_launches = fwd ? out->n_fwd_launches : out->n_bwd_launches;
    first_chunk[ng] = n_chunks;
    first_launch[ng] = n_launches;
    for (int i = 0; i < n; ++i) chunks[n_chunks++] = to_abi(cs[i]);
    if (fwd) {
      n_launches += launches_for(cs, n, qs, nq, causal, launches + n_launches);
    } else {    // backward: one chunk per group, one launch per query chunk that sees it
      for (int q = 0; q < nq; ++q)
        if (visible(qs[q].pos0, qs[q].length, cs[0].pos0, causal)) {
          launches[n_launches].q_chunk = q; launches[n_launches].owner = cs[0].owner;
          launches[n_launches].key_row0 = cs[0].pos0; launches[n_launches].rows = cs[0].length;
          ++n_launches;
        }
    }
    ++ng;
    first_chunk[ng] = n_chunks;
    first_launch[ng] = n_launches;
  };
  if (n_local) add_group(local, n_local, true);
  {
    int i = 0;
    while (i < n_remote) {
      int n = (i == 0 && !n_local) ? 1 : fwd_group_chunks;
      if (n > n_remote - i) n = n_remote - i;
      add_group(remote + i, n, true);
      i += n;
    }
  }
  // backward order: a local chunk first, the remote chunks, the other local chunk last
  if (n_local) add_group(local, 1, false);
  for (int i = 0; i < n_remote; ++i) add_group(remote + i, 1, false);
  for (int i = 1; i < n_local; ++i) add_group(local + i, 1, false);
  // dK/dV partials that land in my heap: (my chunk index, peer) for every peer that needs that chunk
  Ch mine[2];
  const int n_mine = kv_chunks_of(rank, Sk, zigzag, mine);
  for (int ci = 0; ci < n_mine; ++ci)
    for (int peer = 0; peer < world; ++peer) {
      if (peer == rank) continue;
      Ch pn[LWM_RING_MAX_CHUNKS];
      const int np = needed_by(world, peer, Sq, Sk, causal, zigzag, pn);
      bool any = false;
      for (int i = 0; i < np; ++i) any = any || (pn[i].owner == rank && pn[i].index == mine[ci].index);
      if (any) {
        out->incoming[out->n_incoming].chunk_index = mine[ci].index;
        out->incoming[out->n_incoming].peer = peer;
        ++out->n_incoming;
      }
    }
  for (int i = 0; i < n_local; ++i) out->own_computed[out->n_own++] = local[i].index;
  return LWM_OK;
}

// Byte offsets inside every rank's heap payload (two sets, selected by pass parity): identical on all ranks.
extern "C" int lwm_ring_layout(int B, long long Sq, long long Sk, int H, int D, int world, int chunks_per_rank,
                               int op_itemsize, lwm_ring_layout_t* out) {
  if (!out) return lwm_fail(LWM_ERR_ARG, "ring_layout: null out");
  if (B < 1 || Sq < 1 || Sk < 1 || H < 1 || D < 1 || world < 1 || world > LWM_RING_MAX_WORLD || chunks_per_rank < 1 ||
      chunks_per_rank > 2 || Sk % chunks_per_rank || op_itemsize < 1)
    return lwm_fail(LWM_ERR_ARG, "ring_layout: bad arguments");
  const long long row = (long long)H * D;
  long long off = 0;
  out->chunk_rows = Sk / chunks_per_rank;
  out->n_slots = chunks_per_rank * world;
  out->scales = off; off += al256((long long)world * 16);
  out->kg = off; off += al256((long long)B * world * Sk * row * op_itemsize);
  out->vg = off; off += al256((long long)B * world * Sk * row * op_itemsize);
  out->qs = off; off += al256((long long)B * Sq * row * op_itemsize);
  out->lq4 = off; off += al256((long long)B * Sq * row * 4);
  out->lq2 = off; off += al256((long long)B * Sq * row * 2);
  out->slot_bytes = al256((long long)B * out->chunk_rows * row * 4);
  out->lp = off; off += (long long)out->n_slots * 2 * out->slot_bytes;
  out->set_bytes = off;
  out->total = 2 * off;
  return LWM_OK;
}
