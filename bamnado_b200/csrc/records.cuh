// records.cuh — K2: BAM record-boundary discovery over the inflated byte stream of one segment.
//
// Replaces the noodles-bam record iterator the reference drives through `reader.query()`
// (bamnado/src/coverage_analysis.rs:461-466).  BGZF blocks are inflated back to back, so the inflated bytes of a
// segment form one contiguous stream and records that straddle BGZF blocks need no stitching; only the record
// that straddles the *segment* end is carried into the next segment's buffer (carry_copy_kernel).
//
// A record start is only discoverable by chasing `block_size` from a known start, which is serial.  The stream is
// therefore cut into fixed tiles of REC_TILE bytes and solved speculatively:
//   1. rec_guess_walk_kernel  one warp per tile: the lanes test candidate offsets in parallel for a plausible
//                             record header chain; lane 0 then walks the chain to the tile end, storing the
//                             tile-relative start of every record (u16 slots), the record count and the exit
//                             offset (= first record start at or after the tile end).
//   2. rec_fix_scan_kernel    one CTA: tile k+1's entry must equal tile k's exit, seeded by the segment's true
//                             entry; mismatching tiles are re-walked from the true entry until a fixed point, so
//                             the result is the exact chain regardless of how good the guesses were.  Then an
//                             exclusive scan of the per-tile counts gives every tile its first record index.
//   3. rec_compact_kernel     one warp per tile: writes the segment-wide array of record offsets.
// All of it is latency-bound pointer chasing over bytes that K1 just wrote; nothing here is GEMM-shaped.
#pragma once
#include "common.cuh"
#include "kernel_args.h"

namespace bnd {


// Cheap structural test of a BAM record header at `off` (heuristic: only used to pick guesses).
__device__ __forceinline__ bool rec_plausible(const uint8_t *buf, uint64_t off, uint64_t data_end, int32_t n_ref, uint32_t &bs_out) {
  if (off + 36 > data_end) return false;
  const uint8_t *p = buf + off;
  uint32_t bs = ld_u32_unaligned(p);
  if (bs < 32 || bs > REC_MAX_SIZE) return false;
  int32_t ref = (int32_t)ld_u32_unaligned(p + 4), pos = (int32_t)ld_u32_unaligned(p + 8);
  if (ref < -1 || ref >= n_ref || pos < -1) return false;
  uint32_t w3 = ld_u32_unaligned(p + 12), w4 = ld_u32_unaligned(p + 16), l_seq = ld_u32_unaligned(p + 20);
  int32_t nref = (int32_t)ld_u32_unaligned(p + 24), npos = (int32_t)ld_u32_unaligned(p + 28);
  if (nref < -1 || nref >= n_ref || npos < -1) return false;
  uint32_t lrn = w3 & 0xff, ncig = w4 & 0xffff;
  if (lrn == 0) return false;
  uint64_t need = 32ull + lrn + 4ull * ncig + ((uint64_t)l_seq + 1) / 2 + l_seq;
  if (need > bs) return false;
  uint64_t nul = off + 36 + lrn - 1;
  if (nul < data_end && buf[nul] != 0) return false;
  bs_out = bs;
  return true;
}

// A candidate is accepted when it and the next few chained headers all look like records.
__device__ __forceinline__ bool rec_chain_plausible(const uint8_t *buf, uint64_t off, uint64_t data_end, int32_t n_ref) {
  uint64_t p = off;
#pragma unroll 1
  for (int d = 0; d < 4; d++) {
    uint32_t bs;
    if (!rec_plausible(buf, p, data_end, n_ref, bs)) return d > 0 && p + 36 > data_end;
    p += 4ull + bs;
  }
  return true;
}

// Walk the record chain of tile k from `entry`; stores slots/count/exit.  Executed by a single thread.
__device__ __forceinline__ void rec_walk_tile(const RecArgs &a, uint32_t k, uint64_t entry) {
  const uint64_t tile_start = a.tile0 + (uint64_t)k * REC_TILE;
  const uint64_t tile_end = tmin<uint64_t>(tile_start + REC_TILE, a.data_end);
  uint16_t *slots = a.tile_slots + (size_t)k * REC_SLOTS;
  uint32_t n = 0;
  uint64_t p = entry;
  if (entry != REC_NONE && entry >= tile_start) {
#pragma unroll 1
    while (p < tile_end) {
      if (p + 4 > a.data_end) break;  // header of the next record is cut by the segment end
      uint32_t bs = ld_u32_unaligned(a.buf + p);
      if (bs < 32) {
        p = REC_NONE;
        break;
      }
      uint64_t next = p + 4ull + bs;
      if (next > a.data_end) break;  // incomplete record: it is carried into the next segment
      if (n < REC_SLOTS) slots[n] = (uint16_t)(p - tile_start);
      n++;
      p = next;
    }
  }
  a.tile_entry[k] = entry;
  a.tile_exit[k] = p;  // entry < tile_start (record started earlier and is still running) passes through
  a.tile_count[k] = n;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) rec_guess_walk_kernel(RecArgs a) {
  const int lane = (int)(threadIdx.x & 31u);
  const uint32_t k = (uint32_t)((blockIdx.x * (uint64_t)THREADS + threadIdx.x) >> 5);
  if (k >= a.n_tiles) return;  // whole warps leave together
  const uint64_t tile_start = a.tile0 + (uint64_t)k * REC_TILE;
  const uint64_t tile_end = tmin<uint64_t>(tile_start + REC_TILE, a.data_end);
  const uint64_t seg_entry = a.st->entry_off;
  uint64_t guess = REC_NONE;
  if (seg_entry >= tile_start) {
    guess = seg_entry;  // tiles at or before the segment's known first record need no speculation
  } else {
#pragma unroll 1
    for (uint64_t base = tile_start; base < tile_end; base += 32) {
      uint64_t c = base + (uint64_t)lane;
      bool ok = c < tile_end && rec_chain_plausible(a.buf, c, a.data_end, a.n_ref);
      unsigned m = __ballot_sync(0xffffffffu, ok);
      if (m) {
        guess = base + (uint64_t)(__ffs((int)m) - 1);
        break;
      }
    }
  }
  if (lane == 0) rec_walk_tile(a, k, guess);
}

// Single CTA: verify / repair the tile chain, then scan the counts.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) rec_fix_scan_kernel(RecArgs a) {
  __shared__ uint64_t scratch[THREADS / 32 + 1];
  __shared__ uint64_t running;
  const int tid = (int)threadIdx.x;
  const uint64_t seg_entry = a.st->entry_off;
  uint32_t rounds = 0, rewalked = 0;
  // Fixed point without cascades: a tile whose entry differs from its predecessor's exit is re-walked only when the
  // predecessor itself is consistent in this round (its exit will not change under our feet); the thread that owns
  // the head of a run of inconsistent tiles walks the whole run.  A wrong guess therefore costs a couple of rounds
  // instead of rippling through the rest of the segment one tile per round.
  for (;;) {
    int any = 0;
    for (uint32_t k = (uint32_t)tid; k < a.n_tiles; k += THREADS) {
      uint64_t expected = k == 0 ? seg_entry : a.tile_exit[k - 1];
      uint8_t f = a.tile_entry[k] != expected;
      a.tile_flag[k] = f;
      any |= f;
    }
    rounds++;
    if (!__syncthreads_or(any)) break;
    for (uint32_t k = (uint32_t)tid; k < a.n_tiles; k += THREADS) {
      if (a.tile_flag[k] && (k == 0 || !a.tile_flag[k - 1])) {
        // head of a run of inconsistent tiles: walk the whole run from the predecessor's (stable) exit
        uint32_t j = k;
        do {
          rec_walk_tile(a, j, j == 0 ? seg_entry : a.tile_exit[j - 1]);
          rewalked++;
          j++;
        } while (j < a.n_tiles && a.tile_flag[j]);
      }
    }
    __syncthreads();
  }
  // exclusive scan of the per-tile record counts
  if (tid == 0) running = 0;
  __syncthreads();
  for (uint32_t base = 0; base < a.n_tiles; base += THREADS) {
    uint32_t k = base + (uint32_t)tid;
    uint64_t c = k < a.n_tiles ? (uint64_t)a.tile_count[k] : 0ull;
    uint64_t total;
    uint64_t ex = block_exclusive_scan<THREADS, uint64_t>(c, scratch, total);
    if (k < a.n_tiles) a.tile_base[k] = running + ex;
    __syncthreads();
    if (tid == 0) running += total;
    __syncthreads();
  }
  uint32_t rw = warp_sum(rewalked);
  if ((tid & 31) == 0 && rw) atomicAdd(&a.st->rewalked, rw);
  if (tid == 0) {
    SegState *st = a.st;
    uint64_t last_exit = a.n_tiles ? a.tile_exit[a.n_tiles - 1] : seg_entry;
    st->n_records = running;
    st->total_records += running;
    st->fix_rounds += rounds;
    if (last_exit == REC_NONE) {
      if (st->error == SEG_OK) st->error = SEG_ERR_BAD_RECORD;
      st->carry_len = 0;
      st->carry_src = a.data_end;
      st->next_entry_off = a.carry_cap;
    } else {
      uint64_t carry = last_exit < a.data_end ? a.data_end - last_exit : 0;
      // last_exit > data_end cannot happen: a record is only counted when it ends inside the segment
      if (carry > a.carry_cap) {
        if (st->error == SEG_OK) st->error = SEG_ERR_CARRY_TOO_LARGE;
        carry = 0;
      }
      st->carry_len = carry;
      st->carry_src = a.data_end - carry;
      st->next_entry_off = a.carry_cap - carry;
    }
  }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) rec_compact_kernel(RecArgs a) {
  const int lane = (int)(threadIdx.x & 31u);
  const uint32_t k = (uint32_t)((blockIdx.x * (uint64_t)THREADS + threadIdx.x) >> 5);
  if (k >= a.n_tiles) return;
  const uint64_t tile_start = a.tile0 + (uint64_t)k * REC_TILE;
  const uint32_t n = a.tile_count[k];
  const uint64_t base = a.tile_base[k];
  const uint16_t *slots = a.tile_slots + (size_t)k * REC_SLOTS;
  for (uint32_t i = (uint32_t)lane; i < n; i += 32) a.rec_off[base + i] = tile_start + slots[i];
}

// Copies the trailing incomplete record of the previous segment in front of the next segment's inflated bytes and
// opens the next segment's state (each in-flight segment has its own SegState so streams can overlap).
__global__ void carry_copy_kernel(const uint8_t *src_buf, uint8_t *dst_buf, const SegState *prev, SegState *cur, uint64_t carry_cap) {
  const uint64_t n = prev->carry_len;
  const uint64_t src = prev->carry_src, dst = carry_cap - n;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst_buf[dst + i] = src_buf[src + i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    cur->entry_off = prev->next_entry_off;
    cur->next_entry_off = 0;
    cur->n_records = 0;
    cur->carry_len = 0;
    cur->carry_src = 0;
    cur->total_records = prev->total_records;
    cur->last_ref_pos = prev->last_ref_pos;
    cur->error = prev->error;
    cur->fix_rounds = prev->fix_rounds;
    cur->rewalked = prev->rewalked;
    cur->pad = 0;
  }
}

}  // namespace bnd
