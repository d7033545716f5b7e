// Implicit-GEMM "same" convolution on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA,
// with the ConvLSTM gate math fused into the epilogue.
//
//   D[pixel, n] = sum_{src, tap, c}  A_src[pixel + tap_offset, c] * Wp[n, (src, tap, c)]
//
// * M side: one CTA tile = 128 output pixels = a (th x tw) patch of one image (tw*th = 128).
//   For every filter tap the A operand is ONE 4-D TMA box [64 ch, tw, th, 1] of the NHWC source,
//   shifted by the tap offset; out-of-bounds rows/cols are zero-filled by TMA, which IS the
//   reference's zero "same" padding (convlstm.py:12).  No im2col buffer, no torch.cat: x and h
//   are two tensor maps walked in the weight's in-channel order (x first, then h; convlstm.py:17).
// * N side: N_TILE = 4 * CH_TILE columns = gates (i,f,o,g) of CH_TILE hidden channels, packed so
//   the four gates of one channel sit in the same TMEM lane -> the epilogue thread that owns a
//   pixel has i,f,o,g for its channels and applies convlstm.py:21-27 without leaving the SM.
// * K side: blocks of 64 bf16 (=128 B, one SWIZZLE_128B atom row) ordered (src, tap, chunk).
//
// Warp roles (1 CTA/SM, persistent over tiles; with cta_group::2 two CTAs form a pair on one 256-pixel tile):
//   warp 0        : TMA producer (warp-uniform loop, one elected lane issues; smem ring of {A 16 KB, B N_TILE/cta*128 B})
//   warp 1        : tcgen05.mma issuer, leader CTA only (accumulators double-buffered in TMEM: 2 x N_TILE columns)
//   warp 2        : TMEM alloc / dealloc
//   warps 4..     : epilogue -- 16 warps (forward), 8 (gate recompute) or 4 (plain): tcgen05.ld -> gate math ->
//                   smem staging -> TMA tensor stores (64-channel slices) or direct global stores (narrow slices),
//                   overlapped with the next tile's MMAs; the accumulator stage is released right after the last
//                   tcgen05.ld of the tile.
#pragma once
#include "plc_ptx.cuh"

namespace plc {

enum : int { EPI_LSTM_FWD = 0, EPI_LSTM_BWD_GATES = 1, EPI_PLAIN = 2 };

struct ConvTcParams {
  // geometry
  int B, H, W;
  int ksize, pad;
  int tw, th, tw_log2;       // tile = th x tw pixels, tw*th == 128, tw power of two
  int tiles_x, tiles_y;      // per image
  int num_m_tiles, num_n_tiles, num_tiles;
  // exact n / d for n < 2^20 via one 64-bit multiply: q = (n * mul) >> 40, mul = ceil(2^40 / d)  (0 = use '/')
  unsigned long long div_n_tiles, div_tiles_x, div_tiles_y;
  // strided / 3-D convolutions (discriminator; default pipeline only).  H, W above are the OUTPUT grid; the input
  // grid lives in the tensor map, whose elementStrides make a tap box pick every `stride`-th input pixel.  With
  // nd5 the A maps are 5-D [C, W, H, T, B], "images" b = sample * T_out + t_out, and the taps gain a time loop.
  int stride;                // spatial stride (1 or 2); tap box origin = out * stride + tap - pad
  int nd5;                   // 1 = 5-D A tensor maps with a time dimension
  int kt, pad_t, stride_t;   // time taps (1 = none), their padding and stride
  int T_out;                 // output frames per sample (nd5)
  // Tap geometry of the default pipeline: tap (kz, ky, kx), kz < kt, ky < ky_n, kx < kx_n, reads the source at
  //   out * stride + tap0 + k * tap_dir       (plain conv: ky_n = kx_n = ksize, tap0 = -pad, tap_dir = +1).
  // The transposed conv of a STRIDED layer is decomposed by output phase (parity class of the dX pixel): every phase
  // is a stride-1 conv of dZ over the subset of taps k = k0 + j*s whose offsets (phase + pad - k) / s = o0 - j run
  // backwards (tap_dir = -1), and its outputs land on the phase's sub-lattice of dX (o_* below).
  int ky_n, kx_n;
  int tap_x0, tap_y0, tap_t0, tap_dir;
  // EPI_PLAIN output sub-lattice (o_map = 1): tile pixel (b, y, x) -> frame f = sample * o_T + t' * o_st + o_pt (nd5:
  // b = sample * T_out + t'; else f = b), row y * o_s + o_py, column x * o_s + o_px of an [*, o_H, o_W, C] tensor
  int o_map, o_s, o_py, o_px, o_H, o_W, o_st, o_pt, o_T;
  int kc;                    // channels per TMA box: 64, 32 or 16 (narrow sources pack G = 64/kc taps into one K stage)
  int chunks0, chunks1;      // kc-channel chunks of source 0 / source 1
  int num_boxes;             // ksize^2 * (chunks0 + chunks1) boxes of [kc ch x 128 px]
  int num_kb;                // K stages of 64 elements = ceil(num_boxes / (64/kc)); tail boxes repeat the last one
                             // against zero weights
  // patch mode (kc == 64, tile = 16 rows x 8 pixels): ONE haloed [(th+2p) x (tw+2p) px][64 ch] box per (source,
  // 64-channel chunk) serves all ksize^2 taps as shifted UMMA views (profiles/r01_umma_shift_probe.md), instead of
  // ksize^2 separately loaded [128 px][64 ch] tiles: the A feed per K stage drops from 16 KB to 23 KB / 9.
  int patch;                 // 0 = one A tile per K stage (default pipeline)
  int patch_slots;           // resident patches (2 or 3)
  int patch_slot_bytes;      // (th+2p)*(tw+2p)*128 rounded up to 1024
  int b_stages;              // weight-tile ring depth in patch mode
  // LSTM epilogues
  int Ch;                    // hidden channels
  int Cin;                   // EPI_PLAIN: first Cin output columns go to out0, the rest to out1
  const float* bias;         // [4Ch] reference order or nullptr
  const float* c_prev;       // [M, Ch] fp32
  float* c_out;              // [M, Ch] fp32                       (FWD)
  __nv_bfloat16* h_out;      // [M, Ch] bf16                       (FWD)
  __nv_bfloat16* gates_out;  // [M, 4Ch] bf16 or nullptr           (FWD)
  // saved-gates mode (N_TILE = 256): the forward pass keeps the ACTIVATED gates so that BPTT can skip the gate
  // recompute contraction.  Private tile-major layout, 16 bytes per (tile, gate, 8-channel granule, tile row):
  //   gates_saved[(((m_tile * num_n_tiles + n_tile) * 4 + gate) * 8 + granule) * 128 + row]   (uint4 = 8 bf16)
  // -> a forward epilogue warp (32 consecutive tile rows) stores 512 contiguous bytes per instruction, and one
  // 16-channel round of a gate (2 granules) is ONE contiguous 4 KB run for the backward kernel's bulk copies.
  uint4* gates_saved;        // FWD: written; BWD_GATES: read instead of running the mainloop
  const __nv_bfloat16* dh;   // [M, Ch] bf16                       (BWD_GATES)
  const __nv_bfloat16* dh2;  // [M, Ch] bf16 or nullptr, added     (BWD_GATES)
  const float* dc_next;      // [M, Ch] fp32 or nullptr            (BWD_GATES)
  float* dc_prev;            // [M, Ch] fp32                       (BWD_GATES)
  __nv_bfloat16* dz;         // [M, 4Ch] bf16, reference gate order (BWD_GATES)
  __nv_bfloat16* out0;       // [M, Cin] bf16 or nullptr           (PLAIN)
  __nv_bfloat16* out1;       // [M, Ntot-Cin] bf16 or nullptr      (PLAIN)
  int n_total;               // PLAIN: total valid output columns
  const float* plain_bias;   // PLAIN: [n_total] fp32 in packed column order, or nullptr
  int plain_relu;            // PLAIN: apply max(x, plain_slope * x): slope 0 = ReLU, 0.2 = LeakyReLU(0.2)
  float plain_slope;
  int plain_shuffle;         // PLAIN: PixelShuffle(2) store: column n' = sub*(n_total/4) + c -> out0[b, 2y+sub/2, 2x+sub%2, c]
  int plain_f32;             // PLAIN: out0 is an fp32 [M, n_total] tensor (per-thread stores; the last layer of a model keeps
                             // its fp32 accumulator instead of rounding the prediction to bf16)
  int plain_tma;             // PLAIN: outputs leave through smem staging + TMA tensor stores (tmap_o0 / tmap_o1 valid;
                             // needs no shuffle and 64-channel-aligned out0 / out1); 0 = per-thread 16-byte stores
  unsigned long long* prof;  // debug: per-CTA cycle counters [gridDim][16] or nullptr (plc_debug_set_prof)
  int exp_skip_mma;          // EXPERIMENT (PLC_EXP_SKIP_GATE_MMA=1, results are garbage): no operand loads, no MMAs
};

constexpr int kTileM = 128;
constexpr int kBlockK = 64;                       // bf16 elements = 128 bytes
constexpr int kABytes = kTileM * kBlockK * 2;     // 16 KB

// Forward tiles with all 64 channels (N_TILE = 256) leave through shared memory + TMA tensor stores: the per-thread
// NHWC stores (one pixel per lane -> 32 scattered 16/32-byte pieces per instruction) were THE epilogue bottleneck
// (4300 of 10000 cycles per tile); staged, the stores are three bulk copies issued by one thread.
template <int N_TILE, int EPI>
constexpr bool tma_store_epilogue() {
#ifdef PLC_NO_TMA_STORE
  return false;
#else
  return (EPI == 0 /*EPI_LSTM_FWD*/ || EPI == 1 /*EPI_LSTM_BWD_GATES*/) && N_TILE == 256;
#endif
}

template <int N_TILE, int kCta = 1, int EPI = 2>
struct ConvTcCfg {
  static constexpr bool kTmaStore = tma_store_epilogue<N_TILE, EPI>();
  // staging (forward): c' fp32 as two [128 px][32 ch] boxes (2 x 16 KB) + h' bf16 as one [128 px][64 ch] box (16 KB)
  // staging (bwd gates, per 32-channel half): dZ bf16 as four [128 px][32 ch] boxes (4 x 8 KB) + dc_prev fp32 (16 KB)
  // staging (plain): one [128 px][64 ch] bf16 box (16 KB) per 64 output columns
  // (double-buffered up to N_TILE = 128: short-K convs finish a tile faster than its bulk stores drain)
  static constexpr int kPlainBufs = N_TILE <= 128 ? 2 : 1;
  static constexpr int kPlainBufBytes = (N_TILE / 64) * 16384;
  // bwd gates: two operand/output buffers holding one 16-channel round of a tile each (four rounds per 64-channel
  // slice).  The loader warp fills a buffer by TMA with c_prev | dc_next | dh | dh2 of the round; the epilogue overwrites
  // it IN PLACE with dc_prev | dZ_i dZ_f | dZ_o | dZ_g (same bytes per pixel) and the loader warp sends it off as five
  // tensor stores.  Small rounds keep the buffers at 48 KB: the mainloop is weight-feed bound below ~6 weight stages
  // (32-channel rounds = 96 KB left 2 patch slots + 4 stages and the MMA warp waited for weights half of the time).
  static constexpr int kGateRoundCh = 16;
  static constexpr int kGateBufBytes = 128 * kGateRoundCh * 12;     // 24 KB
  static constexpr int kStoreBytes = kTmaStore ? (EPI == 1 /*EPI_LSTM_BWD_GATES*/ ? 2 * kGateBufBytes : 3 * 16384)
                                               : (EPI == 2 /*EPI_PLAIN*/ ? kPlainBufs * kPlainBufBytes : 0);
  // cta_group::2: the CTA pair shares one 256 x N_TILE accumulator tile; each CTA stages its own 128 pixels of A
  // and HALF of the B tile (N_TILE/2 packed-weight rows), so the per-SM operand feed drops from 48 to 32 KB / K-block.
  static constexpr int kBBytes = (N_TILE / kCta) * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kAuxBytes = 4096 + 512;  // bias (<= 1024 floats) + barriers
  static constexpr int kSmemBudget = 227 * 1024 - 1024 /*align slack*/ - kAuxBytes - kStoreBytes;
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  // operand pipeline region: kStages x {A, B} stages, or -- patch mode -- all of it carved into patch slots + weight stages
  static constexpr int kPipeBytes = (kSmemBudget / 1024) * 1024;
  static constexpr int kSmemBytes = kPipeBytes + kStoreBytes + kAuxBytes + 1024;
  // accumulator stages in TMEM: 2 for the wide LSTM tiles (2 x 256 columns); 4 for N_TILE <= 128, where a short-K
  // tile is over in less time than the commit -> epilogue -> release round trip, so the MMA warp must be able to
  // run several tiles ahead of the epilogue
  static constexpr int kAccStages = N_TILE <= 128 ? 4 : 2;
  static constexpr int kAccCols = kAccStages * N_TILE;
  static constexpr int kTmemCols = kAccCols <= 32 ? 32 : kAccCols <= 64 ? 64 : kAccCols <= 128 ? 128
                                   : kAccCols <= 256 ? 256 : 512;
  static_assert(N_TILE % 16 == 0 && N_TILE >= 16 && N_TILE <= 256, "UMMA N constraint for M=128/256");
  static_assert(kCta == 1 || kCta == 2, "cta_group is 1 or 2");
  static_assert(kStages >= 2, "need at least a double buffer");
};

// tile -> (n_tile, image b, patch origin).  With kCta == 2 a "tile" is a PAIR tile: two adjacent 128-pixel m-tiles
// (CTA rank r takes m_tile = 2*mpair + r; an odd tail m-tile decodes to b == B, which TMA zero-fills and the
// epilogue masks).
__device__ __forceinline__ int fast_div(int n, int d, unsigned long long mul) {
  return mul ? static_cast<int>((static_cast<unsigned long long>(n) * mul) >> 40) : n / d;
}
template <int kCta>
__device__ __forceinline__ void decode_tile(const ConvTcParams& p, int tile, int rank, int& n_tile, int& b, int& y0,
                                            int& x0) {
  const int mq = fast_div(tile, p.num_n_tiles, p.div_n_tiles);
  n_tile = tile - mq * p.num_n_tiles;
  const int m_tile = mq * kCta + rank;
  const int r = fast_div(m_tile, p.tiles_x, p.div_tiles_x);
  const int tx = m_tile - r * p.tiles_x;
  b = fast_div(r, p.tiles_y, p.div_tiles_y);
  const int ty = r - b * p.tiles_y;
  y0 = ty * p.th;
  x0 = tx * p.tw;
}

template <int V>
struct IntC { static constexpr int value = V; };

// tensor maps of the gate-gradient epilogue's per-pixel operands (EPI_LSTM_BWD_GATES, N_TILE = 256): 16-channel boxes of
// one 128-pixel tile -- fp32 [16 ch] = 64-byte rows (SWIZZLE_64B), bf16 [16 ch] = 32-byte rows (SWIZZLE_32B)
struct GateMaps { CUtensorMap c_prev, dc_next, dh, dh2; };

// epilogue warps: the LSTM epilogues are MUFU/latency heavy (5 transcendentals per element) and must finish a tile
// faster than the tensor core produces the next one -> two warps per TMEM lane quadrant, alternating 16-channel chunks.
// The forward epilogue works in 8-channel granules with 16 warps (4 per quadrant, 4 per SM sub-partition): it is
// latency bound (TMEM load -> bias -> 5 dependent MUFU/FMA chains per element), so thread-level parallelism is what
// gets a 128 x 64-channel tile through in less than the 9216 cycles the tensor core needs for the next one.
template <int EPI>
constexpr int epi_warps() { return EPI == EPI_PLAIN ? 4 : (EPI == EPI_LSTM_FWD ? 16 : 8); }
template <int EPI>
constexpr int conv_tc_threads() { return 128 + 32 * epi_warps<EPI>(); }

template <int N_TILE, int EPI, int kCta>
__global__ void __launch_bounds__(conv_tc_threads<EPI>(), 1)
conv_igemm_tc_kernel(const ConvTcParams p, const __grid_constant__ CUtensorMap tmap_a0,
                     const __grid_constant__ CUtensorMap tmap_a1, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_o0, const __grid_constant__ CUtensorMap tmap_o1,
                     const __grid_constant__ GateMaps gmaps) {
  using Cfg = ConvTcCfg<N_TILE, kCta, EPI>;
  constexpr int kStages = Cfg::kStages;
  constexpr int CH_TILE = N_TILE / 4;

  constexpr int kEpiWarps = epi_warps<EPI>();
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment (offset arithmetic keeps the pointer in the shared address space)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint8_t* stage_out = smem + Cfg::kPipeBytes;                            // TMA-store staging (1024-aligned)
  float* bias_s = reinterpret_cast<float*>(stage_out + Cfg::kStoreBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + Cfg::kStoreBytes + 4096);
  uint64_t* tmem_full = bars;                      // [2]
  uint64_t* tmem_empty = bars + 4;                 // [kAccStages <= 4]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 8);
  uint64_t* full_bar = bars + 10;                  // [kStages]
  uint64_t* empty_bar = bars + 10 + kStages;       // [kStages]
  // patch mode carves the same regions differently: <= 8 patch slots + <= 8 weight stages (38 barriers of 64)
  uint64_t* patch_full = bars + 10;                // [8]
  uint64_t* patch_empty = bars + 18;               // [8]
  uint64_t* bfull_bar = bars + 26;                 // [8]
  uint64_t* bempty_bar = bars + 34;                // [8]
  uint64_t* gin_full = bars + 44;                  // [4] bwd gates: operand buffer filled by TMA
  uint64_t* gout_ready = bars + 48;                // [4] bwd gates: outputs written, buffer ready for the tensor stores

  // saved-gates form of the gate-gradient kernel: no operand loads, no MMAs; the epilogue takes the activated gates the
  // forward pass stored (bulk copies into the idle operand-pipeline region) instead of the accumulator
  const bool saved = (EPI == EPI_LSTM_BWD_GATES) && p.gates_saved != nullptr;
  // gate-gradient round buffers: two 24 KB in/out buffers behind the operand pipeline (recompute); in the saved form the
  // idle pipeline region holds FOUR 40 KB buffers (24 KB in/out + 16 KB of saved gates) so that the loads run four
  // rounds = one whole tile ahead -- the kernel is HBM-bound and needs the bytes in flight
  const int g_nb = saved ? 4 : 2;
  const uint32_t g_stride = saved ? 40960u : static_cast<uint32_t>(Cfg::kGateBufBytes);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int rank = (kCta == 2) ? static_cast<int>(cluster_ctarank()) : 0;   // 0 = leader (issues the MMAs)
  const int tile0 = blockIdx.x / kCta, tile_step = gridDim.x / kCta;
  const int num_tiles = (kCta == 2) ? ((p.num_m_tiles + 1) / 2) * p.num_n_tiles : p.num_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a0);
    tma_prefetch_desc(&tmap_a1);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    if (p.patch) {
      for (int s = 0; s < 32; ++s) mbar_init(&patch_full[s], 1);   // patch_full/empty[8] + bfull/bempty[8], contiguous
    } else {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
    }
    for (int s = 0; s < Cfg::kAccStages; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps * kCta);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&gin_full[s], 1);
      mbar_init(&gout_ready[s], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kCta>(tmem_ptr_s, Cfg::kTmemCols);
  // Everything above touched only shared memory, TMEM and kernel parameters: under programmatic dependent launch it
  // overlaps the tail of the preceding kernel of the stream.  From here on global memory is read and written.
  pdl_launch_dependents();
  pdl_wait();
  if (EPI != EPI_PLAIN) {
    // forward: sigmoid(x + b) = 0.5 * tanh(0.5 * x + 0.5 * b) + 0.5 -> stage HALF the bias of the i, f, o gates so the
    // bias add folds into the FFMA that scales the pre-activation
    for (int i = threadIdx.x; i < 4 * p.Ch; i += blockDim.x)
      bias_s[i] = (p.bias ? p.bias[i] : 0.f) * ((EPI == EPI_LSTM_FWD && i < 3 * p.Ch) ? 0.5f : 1.f);
  }
  tc_fence_before();
  if constexpr (kCta == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  // The producer and MMA-issuer loops run WARP-UNIFORM (all 32 lanes walk the loop, one elected lane issues):
  // every loop value is then provably uniform, lives in uniform registers, and the per-K-block issue sequence
  // stays far below the 512 cycles the tensor core needs for it (a lane-0-only loop cost ~750 cycles/K-block).
  if (warp == 0) {
    // ===================================================================== TMA producer
    const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b), full_base = smem_u32(full_bar);
    // One specialised copy of the loop per G = 64/kc (boxes per 64-element K stage): the producer sits on the critical
    // path once the kernel is feed-bound, so its per-stage instruction count must stay small (G = 1 is the plain
    // "one box + one weight tile per stage" loop).
    auto produce = [&](auto g_const) {
      constexpr int G = decltype(g_const)::value;
      constexpr int KC = kBlockK / G;
      constexpr uint32_t sub_bytes = kTileM * KC * 2;     // one [128 px][kc ch] box
      uint32_t stage = 0, phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        int n_tile, b, y0, x0;
        decode_tile<kCta>(p, tile, rank, n_tile, b, y0, x0);
        const int n_row = n_tile * N_TILE + rank * (N_TILE / kCta);
        // box iterator over (source, tap = (kz, ky, kx), kc-chunk); the K tail re-loads the last box (zero weights)
        int src = p.chunks0 > 0 ? 0 : 1, kz = 0, ky = 0, kx = 0, ck = 0, bx = 0;
        // strided / 3-D convs: tap origin in INPUT coordinates; image b = sample * T_out + output frame
        const int xs = x0 * p.stride + p.tap_x0, ys = y0 * p.stride + p.tap_y0;
        int smp = b, ts = 0;
        if (p.nd5) { smp = b / p.T_out; ts = (b - smp * p.T_out) * p.stride_t + p.tap_t0; }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          int cch[G], cdx[G], cdy[G], cdt[G], csrc[G];
#pragma unroll
          for (int g = 0; g < G; ++g) {
            cch[g] = ck * KC; cdx[g] = xs + kx * p.tap_dir; cdy[g] = ys + ky * p.tap_dir; cdt[g] = ts + kz * p.tap_dir;
            csrc[g] = src;
            if (bx + 1 < p.num_boxes) {
              ++bx;
              if (++ck == (src ? p.chunks1 : p.chunks0)) {
                ck = 0;
                if (++kx == p.kx_n) {
                  kx = 0;
                  if (++ky == p.ky_n) {
                    ky = 0;
                    if (++kz == p.kt) { kz = 0; src = 1; }
                  }
                }
              }
            }
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            const uint32_t a_dst = a_base + stage * kABytes, b_dst = b_base + stage * Cfg::kBBytes;
            if constexpr (kCta == 1) {
              const uint32_t bar = full_base + stage * 8;
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
              for (int g = 0; g < G; ++g) {
                if (p.nd5)
                  tma_load_5d_s(a_dst + g * sub_bytes, csrc[g] ? &tmap_a1 : &tmap_a0, bar, cch[g], cdx[g], cdy[g], cdt[g],
                                smp);
                else
                  tma_load_4d_s(a_dst + g * sub_bytes, csrc[g] ? &tmap_a1 : &tmap_a0, bar, cch[g], cdx[g], cdy[g], b);
              }
              tma_load_2d_s(b_dst, &tmap_b, bar, kb * kBlockK, n_row);
            } else {
              // both CTAs' bytes are counted on the LEADER's full barrier (the MMA issuer waits there)
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
              const uint32_t bar = mapa_u32(full_base + stage * 8, 0);
#pragma unroll
              for (int g = 0; g < G; ++g) {
                if (p.nd5)
                  tma_load_5d_cg2(a_dst + g * sub_bytes, csrc[g] ? &tmap_a1 : &tmap_a0, bar, cch[g], cdx[g], cdy[g],
                                  cdt[g], smp);
                else
                  tma_load_4d_cg2(a_dst + g * sub_bytes, csrc[g] ? &tmap_a1 : &tmap_a0, bar, cch[g], cdx[g], cdy[g], b);
              }
              tma_load_2d_cg2(b_dst, &tmap_b, bar, kb * kBlockK, n_row);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    };
    // patch mode: per tile and (source, 64-channel chunk) "unit" ONE haloed activation patch, then ksize^2 weight tiles
    auto produce_patch = [&]() {
      const uint32_t pf_base = smem_u32(patch_full), bf_base = smem_u32(bfull_bar);
      const uint32_t wb_base = a_base + p.patch_slots * p.patch_slot_bytes;      // weight ring behind the patch slots
      const int units = p.chunks0 + p.chunks1, taps = p.ksize * p.ksize;
      // weight stages per unit: one per tap with 64-channel boxes; with a single narrow source (kc < 64, one unit)
      // a 64-element K stage holds G = 64/kc consecutive taps
      const int G = kBlockK / p.kc;
      const int kb_per_unit = (taps + G - 1) / G;
      // patch rows are always 128 B (64-channel boxes; a narrow source is zero-filled by TMA beyond its channels)
      const uint32_t patch_tx = (p.tw + 2 * p.pad) * (p.th + 2 * p.pad) * (kBlockK * 2);
      // the patch stream runs ONE unit ahead of the weight stream (also across tile boundaries)
      int tA = tile0, uA = 0, nA = 0, bA = 0, yA = 0, xA = 0;
      uint32_t ps = 0, pphase = 0;
      if (tA < num_tiles) decode_tile<kCta>(p, tA, rank, nA, bA, yA, xA);
      auto issue_patch = [&]() {
        if (tA >= num_tiles) return;
        mbar_wait(&patch_empty[ps], pphase ^ 1);
        if (elect_one()) {
          const int src = uA >= p.chunks0;
          const int ck = src ? uA - p.chunks0 : uA;
          const uint32_t dst = a_base + ps * p.patch_slot_bytes;
          if constexpr (kCta == 1) {
            mbar_arrive_expect_tx(&patch_full[ps], patch_tx);
            tma_load_4d_s(dst, src ? &tmap_a1 : &tmap_a0, pf_base + ps * 8, ck * p.kc, xA - p.pad, yA - p.pad, bA);
          } else {
            if (rank == 0) mbar_arrive_expect_tx(&patch_full[ps], 2 * patch_tx);
            tma_load_4d_cg2(dst, src ? &tmap_a1 : &tmap_a0, mapa_u32(pf_base + ps * 8, 0), ck * p.kc, xA - p.pad,
                            yA - p.pad, bA);
          }
        }
        __syncwarp();
        if (++ps == static_cast<uint32_t>(p.patch_slots)) { ps = 0; pphase ^= 1; }
        if (++uA == units) {
          uA = 0;
          tA += tile_step;
          if (tA < num_tiles) decode_tile<kCta>(p, tA, rank, nA, bA, yA, xA);
        }
      };
      // prologue: a narrow single-unit tile is over in less than one TMA round trip, so its patch stream runs
      // (slots - 1) tiles ahead; 64-channel units last ~9 weight stages: ONE unit ahead (a deeper look-ahead would wait
      // for the slot the MMAs are still reading, and stall the weight stream behind it)
      for (int i = 0; i < (p.kc < 64 ? p.patch_slots - 1 : 1); ++i) issue_patch();
      // With >= 3 slots the next patch is requested before this unit's weights.  With only 2 slots its slot is
      // still being read by the previous unit's MMAs, which are certainly done once the weight ring has wrapped
      // (weight stage t of this unit can only be requested after stage t - b_stages was consumed).
      const int ahead = p.patch_slots >= 3 ? 0 : (p.b_stages < kb_per_unit - 1 ? p.b_stages : kb_per_unit - 1);
      uint32_t bs = 0, bphase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        int n_tile, b, y0, x0;
        decode_tile<kCta>(p, tile, rank, n_tile, b, y0, x0);
        const int n_row = n_tile * N_TILE + rank * (N_TILE / kCta);
        for (int u = 0; u < units; ++u) {
          const int src = u >= p.chunks0;
          const int cks = src ? p.chunks1 : p.chunks0;
          // packed K order: source, tap, chunk (kc == 64); a single narrow unit simply walks its stages in order
          int kb = G > 1 ? 0 : (src ? taps * p.chunks0 + (u - p.chunks0) : u);
          const int kb_step = G > 1 ? 1 : cks;
          for (int t = 0; t < kb_per_unit; ++t, kb += kb_step) {
            if (t == ahead) issue_patch();
            mbar_wait(&bempty_bar[bs], bphase ^ 1);
            if (elect_one()) {
              const uint32_t b_dst = wb_base + bs * Cfg::kBBytes;
              if constexpr (kCta == 1) {
                mbar_arrive_expect_tx(&bfull_bar[bs], Cfg::kBBytes);
                tma_load_2d_s(b_dst, &tmap_b, bf_base + bs * 8, kb * kBlockK, n_row);
              } else {
                if (rank == 0) mbar_arrive_expect_tx(&bfull_bar[bs], 2 * Cfg::kBBytes);
                tma_load_2d_cg2(b_dst, &tmap_b, mapa_u32(bf_base + bs * 8, 0), kb * kBlockK, n_row);
              }
            }
            __syncwarp();
            if (++bs == static_cast<uint32_t>(p.b_stages)) { bs = 0; bphase ^= 1; }
          }
        }
      }
    };
    if (p.exp_skip_mma || saved) {}
    else if (p.patch) produce_patch();
    else if (p.kc == 64) produce(IntC<1>{});
    else if (p.kc == 32) produce(IntC<2>{});
    else produce(IntC<4>{});
  } else if (warp == 1 && rank == 0) {
    // ===================================================================== MMA issuer (leader CTA only)
    constexpr uint32_t idesc = make_idesc_bf16(kTileM * kCta, N_TILE, 0, 0);
    const uint64_t adesc0 = make_smem_desc_kmajor(smem_u32(smem_a), p.kc * 2);   // rows of kc bf16, matching swizzle
    const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_b), 0, 1024);
    // A offset (16-byte units) of the k-th UMMA_K=16 slice of a stage: box (16k / kc), then 32 bytes per slice inside it
    const bool wide = p.kc == 64;
    uint32_t aoff[kBlockK / 16];
#pragma unroll
    for (int k = 0; k < kBlockK / 16; ++k)
      aoff[k] = (((16 * k) / p.kc) * (kTileM * p.kc * 2) + ((16 * k) % p.kc) * 2) >> 4;
    uint32_t stage = 0, phase = 0, bstage = 0, bphase = 0;
    int it = 0;
    long long t_empty = 0, t_full = 0, t_patch = 0, t_mma = 0, t0 = clock64();
    for (int tile = tile0; tile < num_tiles && !saved; tile += tile_step, ++it) {
      const int as = it % Cfg::kAccStages;
      const uint32_t aphase = (it / Cfg::kAccStages) & 1;
      long long ta = (kProfEnabled && p.prof) ? clock64() : 0;
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tc_fence_after();
      if ((kProfEnabled && p.prof)) t_empty += clock64() - ta;
      const uint32_t d_tmem = tmem_base + as * N_TILE;
      if (p.exp_skip_mma) {
        if (elect_one()) {
          if constexpr (kCta == 1) umma_commit<1>(&tmem_full[as]);
          else umma_commit_mc2(&tmem_full[as], 0b11);
        }
        __syncwarp();
        continue;
      }
      if (p.patch) {
        // shifted views: tap (ky, kx) of the tile = the patch read from pixel row ky, pixel kx on (start address
        // + (ky * (tw+2p) + kx) * 128 B), 8-pixel groups (tile rows) one patch row = (tw+2p) * 128 B apart
        const int units = p.chunks0 + p.chunks1, pw = p.tw + 2 * p.pad, taps = p.ksize * p.ksize;
        const int G = kBlockK / p.kc, ksteps = p.kc / 16;           // taps per weight stage, UMMA_K slices per tap
        const int kb_per_unit = (taps + G - 1) / G;
        const uint32_t row_bytes = kBlockK * 2;    // 128-byte SWIZZLE_128B patch rows also for narrow sources
        const uint64_t pdesc0 = make_smem_desc_kmajor_sbo(smem_u32(smem_a), row_bytes, pw * row_bytes);
        const uint64_t wdesc0 = make_smem_desc(smem_u32(smem_a) + p.patch_slots * p.patch_slot_bytes, 0, 1024);
        for (int u = 0; u < units; ++u) {
          long long tb = (kProfEnabled && p.prof) ? clock64() : 0;
          mbar_wait(&patch_full[stage], phase);
          tc_fence_after();
          if ((kProfEnabled && p.prof)) t_patch += clock64() - tb;
          const uint64_t pdesc = pdesc0 + stage * (p.patch_slot_bytes >> 4);
          if (wide) {
            // 64-channel boxes: one weight stage per tap, four UMMA_K slices, fully unrolled issue sequence (this loop
            // must stay well below the 256..512 cycles the tensor core needs per stage)
            for (int ky = 0; ky < p.ksize; ++ky) {
              for (int kx = 0; kx < p.ksize; ++kx) {
                tb = (kProfEnabled && p.prof) ? clock64() : 0;
                mbar_wait(&bfull_bar[bstage], bphase);
                tc_fence_after();
                if ((kProfEnabled && p.prof)) t_full += clock64() - tb;
                const bool last_tap = (ky == p.ksize - 1) && (kx == p.ksize - 1);
                if (elect_one()) {
                  const uint64_t adesc = pdesc + ((ky * pw + kx) * (kBlockK * 2) >> 4);
                  const uint64_t bdesc = wdesc0 + bstage * (Cfg::kBBytes >> 4);
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k)
                    umma_bf16<kCta>(d_tmem, adesc + 2u * k, bdesc + 2 * k, idesc, (u | ky | kx | k) != 0);
                  if constexpr (kCta == 1) {
                    umma_commit<1>(&bempty_bar[bstage]);
                    if (last_tap) umma_commit<1>(&patch_empty[stage]);
                    if (last_tap && u == units - 1) umma_commit<1>(&tmem_full[as]);
                  } else {
                    umma_commit_mc2(&bempty_bar[bstage], 0b11);
                    if (last_tap) umma_commit_mc2(&patch_empty[stage], 0b11);
                    if (last_tap && u == units - 1) umma_commit_mc2(&tmem_full[as], 0b11);
                  }
                }
                __syncwarp();
                if (++bstage == static_cast<uint32_t>(p.b_stages)) { bstage = 0; bphase ^= 1; }
              }
            }
            if (++stage == static_cast<uint32_t>(p.patch_slots)) { stage = 0; phase ^= 1; }
            continue;
          }
          int ky = 0, kx = 0, tap = 0;
          for (int t = 0; t < kb_per_unit; ++t) {
            tb = (kProfEnabled && p.prof) ? clock64() : 0;
            mbar_wait(&bfull_bar[bstage], bphase);
            tc_fence_after();
            if ((kProfEnabled && p.prof)) t_full += clock64() - tb;
            const bool last = t == kb_per_unit - 1;
            const uint64_t bdesc = wdesc0 + bstage * (Cfg::kBBytes >> 4);
            long long tm0 = (kProfEnabled && p.prof) ? clock64() : 0;
            const int ntap = (taps - tap) < G ? (taps - tap) : G;
            if (elect_one()) {       // ONE elected region per stage: all its taps' MMAs and the commits back to back
              int ky2 = ky, kx2 = kx;
              for (int g = 0; g < ntap; ++g) {
                const uint64_t adesc = pdesc + (((ky2 * pw + kx2) * row_bytes) >> 4);
                for (int k = 0; k < ksteps; ++k)
                  umma_bf16<kCta>(d_tmem, adesc + 2u * k, bdesc + 2u * (g * ksteps + k), idesc,
                                  (u | (tap + g) | k) != 0);
                if (++kx2 == p.ksize) { kx2 = 0; ++ky2; }
              }
              if constexpr (kCta == 1) {
                umma_commit<1>(&bempty_bar[bstage]);
                if (last) umma_commit<1>(&patch_empty[stage]);
                if (last && u == units - 1) umma_commit<1>(&tmem_full[as]);
              } else {
                umma_commit_mc2(&bempty_bar[bstage], 0b11);
                if (last) umma_commit_mc2(&patch_empty[stage], 0b11);
                if (last && u == units - 1) umma_commit_mc2(&tmem_full[as], 0b11);
              }
            }
            for (int g = 0; g < ntap; ++g)
              if (++kx == p.ksize) { kx = 0; ++ky; }
            tap += ntap;
            if ((kProfEnabled && p.prof)) t_mma += clock64() - tm0;
            __syncwarp();
            if (++bstage == static_cast<uint32_t>(p.b_stages)) { bstage = 0; bphase ^= 1; }
          }
          if (++stage == static_cast<uint32_t>(p.patch_slots)) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        long long tb = (kProfEnabled && p.prof) ? clock64() : 0;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if ((kProfEnabled && p.prof)) t_full += clock64() - tb;
        if (elect_one()) {
          // descriptor start-address field is in 16-byte units
          const uint64_t adesc = adesc0 + stage * (kABytes >> 4);
          const uint64_t bdesc = bdesc0 + stage * (Cfg::kBBytes >> 4);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // +32 bytes per UMMA_K=16 inside the 128-byte swizzle atom
            umma_bf16<kCta>(d_tmem, adesc + (wide ? 2u * k : aoff[k]), bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if constexpr (kCta == 1) {
            umma_commit<1>(&empty_bar[stage]);
            if (kb == p.num_kb - 1) umma_commit<1>(&tmem_full[as]);
          } else {
            umma_commit_mc2(&empty_bar[stage], 0b11);
            if (kb == p.num_kb - 1) umma_commit_mc2(&tmem_full[as], 0b11);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
    if ((kProfEnabled && p.prof) && lane == 0) {
      p.prof[blockIdx.x * 16 + 0] = clock64() - t0;   // MMA warp: total
      p.prof[blockIdx.x * 16 + 1] = t_empty;          //           waiting for the epilogue to free TMEM
      p.prof[blockIdx.x * 16 + 2] = t_full;           //           waiting for TMA data
      p.prof[blockIdx.x * 16 + 3] = it;               //           tiles
      p.prof[blockIdx.x * 16 + 13] = t_patch;         //           waiting for an activation patch (patch mode)
      p.prof[blockIdx.x * 16 + 14] = t_mma;           //           narrow patch mode: inside the MMA issue loop
    }
  } else if (warp == 3) {
    if constexpr (EPI == EPI_LSTM_BWD_GATES && Cfg::kTmaStore) {
      // ===================================================================== gate-operand loader / output storer
      // Round R = NR * (tile index of this CTA) + (16-channel quarter of the slice) lives in buffer R % g_nb.  Loads run
      // g_nb rounds ahead of the epilogue (2 with the recompute mainloop, 4 in the saved-gates form): buffer R % g_nb is
      // refilled for round R + g_nb as soon as the stores of round R have finished reading it.
      constexpr int RC = Cfg::kGateRoundCh, NR = CH_TILE / RC;
      constexpr uint32_t kF = 128 * RC * 4, kH = 128 * RC * 2;     // bytes of an fp32 / bf16 box
      const uint32_t so = saved ? smem_u32(smem_a) : smem_u32(stage_out), inf_base = smem_u32(gin_full);
      const int my_tiles = tile0 < num_tiles ? (num_tiles - tile0 + tile_step - 1) / tile_step : 0;
      const int rounds = NR * my_tiles;
      const uint32_t tx = kF + kH + (p.dc_next ? kF : 0u) + (p.dh2 ? kH : 0u) + (saved ? 4u * 4096u : 0u);
      // saved gates of a round: [4 gates][2 granules][128 rows][16 B] behind the round's in/out buffer
      auto load_round = [&](int R) {
        if (R >= rounds) return;
        int n_tile, b, y0, x0;
        decode_tile<kCta>(p, tile0 + (R / NR) * tile_step, rank, n_tile, b, y0, x0);
        if (lane == 0) {
          const int bi = R % g_nb;
          const uint32_t buf = so + bi * g_stride, bar = inf_base + bi * 8;
          const int cb = n_tile * CH_TILE + (R % NR) * RC;
          mbar_arrive_expect_tx(&gin_full[bi], tx);
          tma_load_4d_s(buf, &gmaps.c_prev, bar, cb, x0, y0, b);
          if (p.dc_next) tma_load_4d_s(buf + kF, &gmaps.dc_next, bar, cb, x0, y0, b);
          tma_load_4d_s(buf + 2 * kF, &gmaps.dh, bar, cb, x0, y0, b);
          if (p.dh2) tma_load_4d_s(buf + 2 * kF + kH, &gmaps.dh2, bar, cb, x0, y0, b);
          if (saved) {
            const int m_tile = fast_div(tile0 + (R / NR) * tile_step, p.num_n_tiles, p.div_n_tiles) * kCta + rank;
            const uint4* gsrc = p.gates_saved + (static_cast<size_t>(m_tile) * p.num_n_tiles + n_tile) * (4 * 8 * 128) +
                                (R % NR) * (2 * 128);
#pragma unroll
            for (int gate = 0; gate < 4; ++gate)
              bulk_load_1d_s(buf + Cfg::kGateBufBytes + gate * 4096, gsrc + gate * (8 * 128), 4096, bar);
          }
        }
        __syncwarp();
      };
      for (int R = 0; R < g_nb; ++R) load_round(R);
      for (int R = 0; R < rounds; ++R) {
        mbar_wait(&gout_ready[R % g_nb], (R / g_nb) & 1);
        int n_tile, b, y0, x0;
        decode_tile<kCta>(p, tile0 + (R / NR) * tile_step, rank, n_tile, b, y0, x0);
        if (lane == 0) {   // OOB rows / images (ragged tiles, odd tail pair) are clipped by TMA
          const uint32_t buf = so + (R % g_nb) * g_stride;
          const int cb = n_tile * CH_TILE + (R % NR) * RC;
          tma_store_4d(&tmap_o1, buf, cb, x0, y0, b);                                    // dc_prev
#pragma unroll
          for (int gate = 0; gate < 4; ++gate)                                           // dZ, reference gate order
            tma_store_4d(&tmap_o0, buf + kF + gate * kH, gate * p.Ch + cb, x0, y0, b);
          tma_store_commit();
          tma_store_wait_read();
        }
        __syncwarp();
        load_round(R + g_nb);
      }
      if (lane == 0) tma_store_wait_all();   // bulk stores complete before the CTA retires
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue
    const int q = warp & 3;             // TMEM lane quadrant == warp_idx % 4
    const int half = (warp - 4) >> 2;   // which of the kEpiWarps/4 warps of this quadrant (a.k.a. wq)
    constexpr int kChunkStep = kEpiWarps / 4;
    const int row = q * 32 + lane;
    const int ty = row >> p.tw_log2;
    const int tx = row & (p.tw - 1);
    int it = 0;
    long long epi_t0 = 0, pa_wait = 0, pa_busy = 0, pa_barA = 0, pa_ld = 0, pa_math = 0, pa_barB = 0, pa_pre = 0, pa_iss = 0, pa_top = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
      const long long tTop = ((kProfEnabled && p.prof) && warp == 4 && lane == 0) ? clock64() : 0;
      const int as = it % Cfg::kAccStages;
      const uint32_t aphase = (it / Cfg::kAccStages) & 1;
      int n_tile, b, y0, x0;
      decode_tile<kCta>(p, tile, rank, n_tile, b, y0, x0);
      const int y = y0 + ty, x = x0 + tx;
      const bool valid = (y < p.H) && (x < p.W) && (b < p.B);
      const size_t pix = (static_cast<size_t>(b) * p.H + y) * p.W + x;
      // hand the accumulator stage back to the MMA warp as soon as this warp's last tcgen05.ld has landed
      bool released = false;
      auto release = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCta == 1) mbar_arrive(&tmem_empty[as]);
          else mbar_arrive_cluster(&tmem_empty[as], 0);   // the leader's barrier gates the next MMA into this stage
        }
        released = true;
      };
      constexpr int kLstmChunks = CH_TILE / 16;
      constexpr int kMaxMine = (kLstmChunks + kChunkStep - 1) / kChunkStep;
      constexpr int kGran = CH_TILE / 8;                                   // forward: 8-channel granules
      constexpr int kMaxGran = (kGran + kChunkStep - 1) / kChunkStep;
      // forward: fetch this thread's c_prev BEFORE waiting for the accumulator -> DRAM latency hides under the MMAs
      float cpre[EPI == EPI_LSTM_FWD ? kMaxGran : 1][8];
      if constexpr (EPI == EPI_LSTM_FWD) {
#pragma unroll
        for (int m = 0; m < kMaxGran; ++m) {
          const int g = half + m * kChunkStep;
          if (valid && g < kGran) {
            if (p.c_prev) {
              const float4* src = reinterpret_cast<const float4*>(p.c_prev + pix * p.Ch + n_tile * CH_TILE + g * 8);
              const float4 t0 = __ldg(src), t1 = __ldg(src + 1);
              cpre[m][0] = t0.x; cpre[m][1] = t0.y; cpre[m][2] = t0.z; cpre[m][3] = t0.w;
              cpre[m][4] = t1.x; cpre[m][5] = t1.y; cpre[m][6] = t1.z; cpre[m][7] = t1.w;
            } else {                 // zero initial state (generator.py:156-160): c_prev == 0
#pragma unroll
              for (int e = 0; e < 8; ++e) cpre[m][e] = 0.f;
            }
          }
        }
      }
      long long te = ((kProfEnabled && p.prof) && warp == 4) ? clock64() : 0;
      if ((kProfEnabled && p.prof) && warp == 4 && lane == 0) pa_pre += te - tTop;     // decode + operand prefetch issue
      if (!saved) {
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
      }
      if ((kProfEnabled && p.prof) && warp == 4 && lane == 0) {
        const long long now = clock64();
        pa_wait += now - te;                    // epilogue warp 4: waiting for an accumulator
        if (it > 0) pa_busy += te - epi_t0;     //                  busy (previous tile's work)
        epi_t0 = now;
      }
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * N_TILE;

      if constexpr (EPI == EPI_LSTM_FWD) {
        const int ch0 = n_tile * CH_TILE;
        const bool issuer = (warp == 4) && (lane == 0);
        const bool pw4 = (kProfEnabled && p.prof) && warp == 4 && lane == 0;
        long long tA = pw4 ? clock64() : 0;
        if constexpr (Cfg::kTmaStore) {
          // staging buffer free again?  (the issuer's previous bulk stores have finished reading it)
          if (issuer) tma_store_wait_read();
          named_bar_sync(1, 32 * kEpiWarps);
        }
        if (pw4) pa_barA += clock64() - tA;   // barrier A (staging free)
#pragma unroll
        for (int m = 0; m < kMaxGran; ++m) {
          const int g = half + m * kChunkStep;
          if (g >= kGran) break;
          uint32_t vi[8], vf[8], vo[8], vg[8];
          tmem_ld8(t_acc + 0 * CH_TILE + g * 8, vi);
          tmem_ld8(t_acc + 1 * CH_TILE + g * 8, vf);
          tmem_ld8(t_acc + 2 * CH_TILE + g * 8, vo);
          tmem_ld8(t_acc + 3 * CH_TILE + g * 8, vg);
          const int chb = ch0 + g * 8;
          // (half-)bias of this granule: smem reads issued before the TMEM wait so both latencies overlap
          float bi[8], bf[8], bo[8], bg[8];
          {
            const float4* s0 = reinterpret_cast<const float4*>(bias_s + 0 * p.Ch + chb);
            const float4* s1 = reinterpret_cast<const float4*>(bias_s + 1 * p.Ch + chb);
            const float4* s2 = reinterpret_cast<const float4*>(bias_s + 2 * p.Ch + chb);
            const float4* s3 = reinterpret_cast<const float4*>(bias_s + 3 * p.Ch + chb);
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              const float4 a = s0[v], b4 = s1[v], c4 = s2[v], d4 = s3[v];
              bi[4 * v] = a.x; bi[4 * v + 1] = a.y; bi[4 * v + 2] = a.z; bi[4 * v + 3] = a.w;
              bf[4 * v] = b4.x; bf[4 * v + 1] = b4.y; bf[4 * v + 2] = b4.z; bf[4 * v + 3] = b4.w;
              bo[4 * v] = c4.x; bo[4 * v + 1] = c4.y; bo[4 * v + 2] = c4.z; bo[4 * v + 3] = c4.w;
              bg[4 * v] = d4.x; bg[4 * v + 1] = d4.y; bg[4 * v + 2] = d4.z; bg[4 * v + 3] = d4.w;
            }
          }
          long long tL = pw4 ? clock64() : 0;
          tmem_ld_wait();
          if (pw4) pa_ld += clock64() - tL;   // tcgen05.wait::ld
          if (g + kChunkStep >= kGran) release();
          tL = pw4 ? clock64() : 0;
          if (valid) {
            const size_t off = pix * p.Ch + chb;
            float cn[8];
            uint32_t hp[4];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              float hv[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int jj = j + u;
                // convlstm.py:21-24 (bias already halved for the sigmoid gates)
                const float ig = fmaf(tanh_fast(fmaf(__uint_as_float(vi[jj]), 0.5f, bi[jj])), 0.5f, 0.5f);
                const float fg = fmaf(tanh_fast(fmaf(__uint_as_float(vf[jj]), 0.5f, bf[jj])), 0.5f, 0.5f);
                const float og = fmaf(tanh_fast(fmaf(__uint_as_float(vo[jj]), 0.5f, bo[jj])), 0.5f, 0.5f);
                const float gt = tanh_fast(__uint_as_float(vg[jj]) + bg[jj]);
                const float c2 = fmaf(fg, cpre[m][jj], ig * gt);       // convlstm.py:26
                cn[jj] = c2;
                hv[u] = og * tanh_fast(c2);                            // convlstm.py:27
                vi[jj] = __float_as_uint(ig); vf[jj] = __float_as_uint(fg);
                vo[jj] = __float_as_uint(og); vg[jj] = __float_as_uint(gt);
              }
              hp[j >> 1] = pack_bf16x2(hv[0], hv[1]);
            }
            if constexpr (Cfg::kTmaStore) {
              // SWIZZLE_128B staging: 16-byte chunk index XOR (row % 8) -> conflict-free st.shared.v4
              const uint32_t sw = row & 7;
              const uint32_t cb = smem_u32(stage_out) + (g >> 2) * 16384 + row * 128;
              const uint32_t j0 = (g & 3) * 2;
              st_shared_v4(cb + ((j0 ^ sw) << 4), __float_as_uint(cn[0]), __float_as_uint(cn[1]),
                           __float_as_uint(cn[2]), __float_as_uint(cn[3]));
              st_shared_v4(cb + (((j0 + 1) ^ sw) << 4), __float_as_uint(cn[4]), __float_as_uint(cn[5]),
                           __float_as_uint(cn[6]), __float_as_uint(cn[7]));
              st_shared_v4(smem_u32(stage_out) + 2 * 16384 + row * 128 + ((static_cast<uint32_t>(g) ^ sw) << 4), hp[0],
                           hp[1], hp[2], hp[3]);
            } else {
              float4* cdst = reinterpret_cast<float4*>(p.c_out + off);
              cdst[0] = make_float4(cn[0], cn[1], cn[2], cn[3]);
              cdst[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
              *reinterpret_cast<uint4*>(p.h_out + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
            }
            if (p.gates_saved) {
              if constexpr (N_TILE == 256) {
                const int m_tile = fast_div(tile, p.num_n_tiles, p.div_n_tiles) * kCta + rank;
                uint4* gs = p.gates_saved + (static_cast<size_t>(m_tile) * p.num_n_tiles + n_tile) * (4 * 8 * 128) +
                            g * 128 + row;
#pragma unroll
                for (int gate = 0; gate < 4; ++gate) {
                  const uint32_t* src = gate == 0 ? vi : gate == 1 ? vf : gate == 2 ? vo : vg;
                  // streaming store (evict-first): written once, read once a whole rollout later -- it must not push the
                  // activation patches the mainloop re-reads out of L2
                  __stcs(gs + gate * (8 * 128),
                         make_uint4(pack_bf16x2(__uint_as_float(src[0]), __uint_as_float(src[1])),
                                    pack_bf16x2(__uint_as_float(src[2]), __uint_as_float(src[3])),
                                    pack_bf16x2(__uint_as_float(src[4]), __uint_as_float(src[5])),
                                    pack_bf16x2(__uint_as_float(src[6]), __uint_as_float(src[7]))));
                }
              }
            }
            if (p.gates_out) {
              __nv_bfloat16* gbase = p.gates_out + pix * (4 * p.Ch) + chb;
#pragma unroll
              for (int gate = 0; gate < 4; ++gate) {
                const uint32_t* src = gate == 0 ? vi : gate == 1 ? vf : gate == 2 ? vo : vg;
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(src[0]), __uint_as_float(src[1]));
                o.y = pack_bf16x2(__uint_as_float(src[2]), __uint_as_float(src[3]));
                o.z = pack_bf16x2(__uint_as_float(src[4]), __uint_as_float(src[5]));
                o.w = pack_bf16x2(__uint_as_float(src[6]), __uint_as_float(src[7]));
                *reinterpret_cast<uint4*>(gbase + gate * p.Ch) = o;
              }
            }
          }
          if (pw4) pa_math += clock64() - tL;   // gate math + st.shared
        }
        long long tB = pw4 ? clock64() : 0;
        if constexpr (Cfg::kTmaStore) {
          fence_proxy_async_smem();               // my st.shared -> visible to the TMA (async proxy)
          named_bar_sync(1, 32 * kEpiWarps);
          if (pw4) pa_barB += clock64() - tB;   // fence + barrier B
          if (issuer) {                           // OOB rows / images (ragged tiles, odd tail pair) are clipped by TMA
            const long long tS = (kProfEnabled && p.prof) ? clock64() : 0;
            const uint32_t so = smem_u32(stage_out);
            tma_store_4d(&tmap_o0, so, ch0, x0, y0, b);
            tma_store_4d(&tmap_o0, so + 16384, ch0 + 32, x0, y0, b);
            tma_store_4d(&tmap_o1, so + 2 * 16384, ch0, x0, y0, b);
            tma_store_commit();
            if ((kProfEnabled && p.prof)) pa_iss += clock64() - tS;
          }
        }
      } else if constexpr (EPI == EPI_LSTM_BWD_GATES && Cfg::kTmaStore) {
        // ---- gate recompute + dZ / dc_prev, 8-channel granules.  Four 16-channel rounds per tile.
        // The per-pixel operands arrive by TMA in the round's buffer (loader warp) and the outputs
        // overwrite them in place: dc_prev over c_prev, dZ_o over dh, dZ_g over dh2 (same thread, same address), dZ_i and
        // dZ_f over dc_next -- the only aliasing across threads, hence dc_next is read first and barrier X follows.
        // No per-thread global access is left in this epilogue; ragged tiles compute on TMA's zero fill and are clipped
        // by the tensor stores.
        const int ch0 = n_tile * CH_TILE;
        constexpr int RC = Cfg::kGateRoundCh, NR = CH_TILE / RC, GR = RC / 8;
        constexpr int WQ = kEpiWarps / 4;          // warps per TMEM lane quadrant
        constexpr int GPR = GR / WQ;               // granules per round per warp
        constexpr uint32_t kF = 128 * RC * 4, kH = 128 * RC * 2;     // bytes of an fp32 / bf16 box
        static_assert(CH_TILE == 64 && RC == 16 && GPR >= 1 && GPR * WQ == GR, "gate epilogue: 16-channel rounds, 8 warps");
        const uint32_t sw64 = (row >> 1) & 3, sw32 = (row >> 2) & 1;
        const bool pw4 = (kProfEnabled && p.prof) && warp == 4 && lane == 0;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const int gR = it * NR + r, gbi = gR % g_nb;       // global round index of this CTA, its buffer
          const uint32_t buf = (saved ? smem_u32(smem_a) : smem_u32(stage_out)) + gbi * g_stride;
          const uint32_t crow = buf + row * 64;                  // c_prev -> dc_prev   [128 px][16 ch] fp32, SWIZZLE_64B
          const uint32_t nrow = buf + kF + row * 64;             // dc_next             (same layout)
          const uint32_t zrow = buf + kF + row * 32;             // dZ_i, dZ_f, dZ_o (over dh), dZ_g (over dh2): four
                                                                 // [128 px][16 ch] bf16 boxes, 32-byte rows, SWIZZLE_32B
          long long tA = pw4 ? clock64() : 0;
          mbar_wait(&gin_full[gbi], (gR / g_nb) & 1);
          if (pw4) pa_barA += clock64() - tA;                    // waiting for the round's operands
          uint4 dn[GPR][2];
#pragma unroll
          for (int j = 0; j < GPR; ++j) {
            const uint32_t gl = half * GPR + j;
            if (p.dc_next) {
              dn[j][0] = ld_shared_v4(nrow + (((gl * 2) ^ sw64) << 4));
              dn[j][1] = ld_shared_v4(nrow + (((gl * 2 + 1) ^ sw64) << 4));
            } else {
              dn[j][0] = make_uint4(0u, 0u, 0u, 0u); dn[j][1] = dn[j][0];
            }
          }
          tA = pw4 ? clock64() : 0;
          named_bar_sync(1, 32 * kEpiWarps);                     // X: every dc_next read precedes every dZ_i / dZ_f write
          if (pw4) pa_barB += clock64() - tA;
#pragma unroll
          for (int j = 0; j < GPR; ++j) {
            constexpr int kLast = NR * GPR - 1;
            const int idx = r * GPR + j;
            const uint32_t gl = half * GPR + j;     // granule within the round
            const int g = r * GR + gl;              // granule within the 64-channel slice: 0..7
            uint32_t vi[8], vf[8], vo[8], vg[8];
            uint4 sg[4];                            // saved mode: activated gates i, f, o, g of this granule (8 x bf16 each)
            if (!saved) {
              tmem_ld8(t_acc + 0 * CH_TILE + g * 8, vi);
              tmem_ld8(t_acc + 1 * CH_TILE + g * 8, vf);
              tmem_ld8(t_acc + 2 * CH_TILE + g * 8, vo);
              tmem_ld8(t_acc + 3 * CH_TILE + g * 8, vg);
            } else {
              const uint32_t ga = buf + Cfg::kGateBufBytes + gl * 2048 + row * 16;
#pragma unroll
              for (int gate = 0; gate < 4; ++gate) sg[gate] = ld_shared_v4(ga + gate * 4096);
            }
            const int chb = ch0 + g * 8;
            float bi[8], bf[8], bo[8], bg[8];
            if (!saved) {
              const float4* s0 = reinterpret_cast<const float4*>(bias_s + 0 * p.Ch + chb);
              const float4* s1 = reinterpret_cast<const float4*>(bias_s + 1 * p.Ch + chb);
              const float4* s2 = reinterpret_cast<const float4*>(bias_s + 2 * p.Ch + chb);
              const float4* s3 = reinterpret_cast<const float4*>(bias_s + 3 * p.Ch + chb);
#pragma unroll
              for (int v = 0; v < 2; ++v) {
                const float4 a = s0[v], b4 = s1[v], c4 = s2[v], d4 = s3[v];
                bi[4 * v] = a.x; bi[4 * v + 1] = a.y; bi[4 * v + 2] = a.z; bi[4 * v + 3] = a.w;
                bf[4 * v] = b4.x; bf[4 * v + 1] = b4.y; bf[4 * v + 2] = b4.z; bf[4 * v + 3] = b4.w;
                bo[4 * v] = c4.x; bo[4 * v + 1] = c4.y; bo[4 * v + 2] = c4.z; bo[4 * v + 3] = c4.w;
                bg[4 * v] = d4.x; bg[4 * v + 1] = d4.y; bg[4 * v + 2] = d4.z; bg[4 * v + 3] = d4.w;
              }
            }
            const uint32_t ca0 = crow + (((gl * 2) ^ sw64) << 4), ca1 = crow + (((gl * 2 + 1) ^ sw64) << 4);
            const uint32_t za = zrow + ((gl ^ sw32) << 4);
            const uint4 c0 = ld_shared_v4(ca0), c1 = ld_shared_v4(ca1);
            const uint4 h1 = ld_shared_v4(za + 2 * kH);
            const uint4 h2 = p.dh2 ? ld_shared_v4(za + 3 * kH) : make_uint4(0u, 0u, 0u, 0u);
            long long tL = pw4 ? clock64() : 0;
            if (!saved) tmem_ld_wait();
            if (pw4) pa_ld += clock64() - tL;
            if (idx == kLast && !saved) release();
            tL = pw4 ? clock64() : 0;
            {
              const float cp[8] = {__uint_as_float(c0.x), __uint_as_float(c0.y), __uint_as_float(c0.z), __uint_as_float(c0.w),
                                   __uint_as_float(c1.x), __uint_as_float(c1.y), __uint_as_float(c1.z), __uint_as_float(c1.w)};
              const float dcn[8] = {__uint_as_float(dn[j][0].x), __uint_as_float(dn[j][0].y), __uint_as_float(dn[j][0].z),
                                    __uint_as_float(dn[j][0].w), __uint_as_float(dn[j][1].x), __uint_as_float(dn[j][1].y),
                                    __uint_as_float(dn[j][1].z), __uint_as_float(dn[j][1].w)};
              const uint32_t w1[4] = {h1.x, h1.y, h1.z, h1.w};
              const uint32_t w2[4] = {h2.x, h2.y, h2.z, h2.w};
              float dcp[8];
              uint32_t zi[4], zf[4], zo[4], zg[4];
#pragma unroll
              for (int e = 0; e < 8; e += 2) {
                const __nv_bfloat162 t1 = *reinterpret_cast<const __nv_bfloat162*>(&w1[e >> 1]);
                const __nv_bfloat162 t2 = *reinterpret_cast<const __nv_bfloat162*>(&w2[e >> 1]);
                const float dhp[2] = {__low2float(t1) + __low2float(t2), __high2float(t1) + __high2float(t2)};
                float di_[2], df_[2], do_[2], dg_[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int jj = e + u;
                  float ig, fg, og, gt;
                  if (!saved) {
                    ig = sigmoid_fast(__uint_as_float(vi[jj]) + bi[jj]);
                    fg = sigmoid_fast(__uint_as_float(vf[jj]) + bf[jj]);
                    og = sigmoid_fast(__uint_as_float(vo[jj]) + bo[jj]);
                    gt = tanh_fast(__uint_as_float(vg[jj]) + bg[jj]);
                  } else {                       // bf16 -> fp32: the stored half is the high half of the float
                    const uint32_t wi = (&sg[0].x)[e >> 1], wf = (&sg[1].x)[e >> 1];
                    const uint32_t wo = (&sg[2].x)[e >> 1], wg = (&sg[3].x)[e >> 1];
                    ig = __uint_as_float(u ? (wi & 0xffff0000u) : (wi << 16));
                    fg = __uint_as_float(u ? (wf & 0xffff0000u) : (wf << 16));
                    og = __uint_as_float(u ? (wo & 0xffff0000u) : (wo << 16));
                    gt = __uint_as_float(u ? (wg & 0xffff0000u) : (wg << 16));
                  }
                  const float c2 = fmaf(fg, cp[jj], ig * gt);
                  const float tc = tanh_fast(c2);
                  const float dh_ = dhp[u];
                  const float dc = fmaf(dh_ * og, 1.f - tc * tc, dcn[jj]);
                  dcp[jj] = dc * fg;
                  di_[u] = dc * gt * ig * (1.f - ig);
                  df_[u] = dc * cp[jj] * fg * (1.f - fg);
                  do_[u] = dh_ * tc * og * (1.f - og);
                  dg_[u] = dc * ig * (1.f - gt * gt);
                }
                zi[e >> 1] = pack_bf16x2(di_[0], di_[1]);
                zf[e >> 1] = pack_bf16x2(df_[0], df_[1]);
                zo[e >> 1] = pack_bf16x2(do_[0], do_[1]);
                zg[e >> 1] = pack_bf16x2(dg_[0], dg_[1]);
              }
              st_shared_v4(za + 0 * kH, zi[0], zi[1], zi[2], zi[3]);
              st_shared_v4(za + 1 * kH, zf[0], zf[1], zf[2], zf[3]);
              st_shared_v4(za + 2 * kH, zo[0], zo[1], zo[2], zo[3]);
              st_shared_v4(za + 3 * kH, zg[0], zg[1], zg[2], zg[3]);
              st_shared_v4(ca0, __float_as_uint(dcp[0]), __float_as_uint(dcp[1]), __float_as_uint(dcp[2]),
                           __float_as_uint(dcp[3]));
              st_shared_v4(ca1, __float_as_uint(dcp[4]), __float_as_uint(dcp[5]), __float_as_uint(dcp[6]),
                           __float_as_uint(dcp[7]));
            }
            if (pw4) pa_math += clock64() - tL;
          }
          fence_proxy_async_smem();                 // my st.shared -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&gout_ready[gbi]);
        }
      } else if constexpr (EPI == EPI_LSTM_BWD_GATES) {
        // ---- direct-store variant (channel slices narrower than 64): 16-channel chunks
        const int ch0 = n_tile * CH_TILE;
#pragma unroll
        for (int m = 0; m < kMaxMine; ++m) {
          const int cc = half + m * kChunkStep;
          if (cc >= kLstmChunks) break;
          uint32_t vi[16], vf[16], vo[16], vg[16];
          tmem_ld16(t_acc + 0 * CH_TILE + cc * 16, vi);
          tmem_ld16(t_acc + 1 * CH_TILE + cc * 16, vf);
          tmem_ld16(t_acc + 2 * CH_TILE + cc * 16, vo);
          tmem_ld16(t_acc + 3 * CH_TILE + cc * 16, vg);
          tmem_ld_wait();
          if (cc + kChunkStep >= kLstmChunks) release();
          if (valid) {
            const int chb = ch0 + cc * 16;
            const size_t off = pix * p.Ch + chb;
            {  // + bias (reference order: gate*Ch + channel), 16-byte smem broadcasts
#pragma unroll
              for (int v4 = 0; v4 < 4; ++v4) {
                const float4 bi = *reinterpret_cast<const float4*>(bias_s + 0 * p.Ch + chb + 4 * v4);
                const float4 bf = *reinterpret_cast<const float4*>(bias_s + 1 * p.Ch + chb + 4 * v4);
                const float4 bo = *reinterpret_cast<const float4*>(bias_s + 2 * p.Ch + chb + 4 * v4);
                const float4 bg = *reinterpret_cast<const float4*>(bias_s + 3 * p.Ch + chb + 4 * v4);
                const float bia[4] = {bi.x, bi.y, bi.z, bi.w}, bfa[4] = {bf.x, bf.y, bf.z, bf.w};
                const float boa[4] = {bo.x, bo.y, bo.z, bo.w}, bga[4] = {bg.x, bg.y, bg.z, bg.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  vi[4 * v4 + e] = __float_as_uint(__uint_as_float(vi[4 * v4 + e]) + bia[e]);
                  vf[4 * v4 + e] = __float_as_uint(__uint_as_float(vf[4 * v4 + e]) + bfa[e]);
                  vo[4 * v4 + e] = __float_as_uint(__uint_as_float(vo[4 * v4 + e]) + boa[e]);
                  vg[4 * v4 + e] = __float_as_uint(__uint_as_float(vg[4 * v4 + e]) + bga[e]);
                }
              }
            }
            float cp[16];
            {
              const float4* src = reinterpret_cast<const float4*>(p.c_prev + off);
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                float4 t = __ldg(src + v);
                cp[4 * v + 0] = t.x; cp[4 * v + 1] = t.y; cp[4 * v + 2] = t.z; cp[4 * v + 3] = t.w;
              }
            }
            {
              float dhv[16], dcn[16], dcp[16];
              {
                const uint4* s = reinterpret_cast<const uint4*>(p.dh + off);
                uint4 a = __ldg(s), bq = __ldg(s + 1);
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
                  dhv[2 * j] = __low2float(t); dhv[2 * j + 1] = __high2float(t);
                }
                if (p.dh2) {
                  const uint4* s2 = reinterpret_cast<const uint4*>(p.dh2 + off);
                  uint4 a2 = __ldg(s2), b2 = __ldg(s2 + 1);
                  const uint32_t w2[8] = {a2.x, a2.y, a2.z, a2.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w2[j]);
                    dhv[2 * j] += __low2float(t); dhv[2 * j + 1] += __high2float(t);
                  }
                }
              }
              if (p.dc_next) {
                const float4* s = reinterpret_cast<const float4*>(p.dc_next + off);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                  float4 t = __ldg(s + v);
                  dcn[4 * v] = t.x; dcn[4 * v + 1] = t.y; dcn[4 * v + 2] = t.z; dcn[4 * v + 3] = t.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) dcn[j] = 0.f;
              }
              uint32_t zi[8], zf[8], zo[8], zg[8];
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                float di_[2], df_[2], do_[2], dg_[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int jj = j + u;
                  const float ig = sigmoid_fast(__uint_as_float(vi[jj]));
                  const float fg = sigmoid_fast(__uint_as_float(vf[jj]));
                  const float og = sigmoid_fast(__uint_as_float(vo[jj]));
                  const float gt = tanh_fast(__uint_as_float(vg[jj]));
                  const float c2 = fmaf(fg, cp[jj], ig * gt);
                  const float tc = tanh_fast(c2);
                  const float dh_ = dhv[jj];
                  const float dc = fmaf(dh_ * og, 1.f - tc * tc, dcn[jj]);
                  dcp[jj] = dc * fg;
                  di_[u] = dc * gt * ig * (1.f - ig);
                  df_[u] = dc * cp[jj] * fg * (1.f - fg);
                  do_[u] = dh_ * tc * og * (1.f - og);
                  dg_[u] = dc * ig * (1.f - gt * gt);
                }
                zi[j >> 1] = pack_bf16x2(di_[0], di_[1]);
                zf[j >> 1] = pack_bf16x2(df_[0], df_[1]);
                zo[j >> 1] = pack_bf16x2(do_[0], do_[1]);
                zg[j >> 1] = pack_bf16x2(dg_[0], dg_[1]);
              }
              float4* dcd = reinterpret_cast<float4*>(p.dc_prev + off);
#pragma unroll
              for (int v = 0; v < 4; ++v)
                dcd[v] = make_float4(dcp[4 * v], dcp[4 * v + 1], dcp[4 * v + 2], dcp[4 * v + 3]);
              __nv_bfloat16* zb = p.dz + pix * (4 * p.Ch) + chb;
              uint4* d0 = reinterpret_cast<uint4*>(zb + 0 * p.Ch);
              uint4* d1 = reinterpret_cast<uint4*>(zb + 1 * p.Ch);
              uint4* d2 = reinterpret_cast<uint4*>(zb + 2 * p.Ch);
              uint4* d3 = reinterpret_cast<uint4*>(zb + 3 * p.Ch);
              d0[0] = make_uint4(zi[0], zi[1], zi[2], zi[3]); d0[1] = make_uint4(zi[4], zi[5], zi[6], zi[7]);
              d1[0] = make_uint4(zf[0], zf[1], zf[2], zf[3]); d1[1] = make_uint4(zf[4], zf[5], zf[6], zf[7]);
              d2[0] = make_uint4(zo[0], zo[1], zo[2], zo[3]); d2[1] = make_uint4(zo[4], zo[5], zo[6], zo[7]);
              d3[0] = make_uint4(zg[0], zg[1], zg[2], zg[3]); d3[1] = make_uint4(zg[4], zg[5], zg[6], zg[7]);
            }
          }
        }
      } else if (p.plain_tma) {
        // EPI_PLAIN through shared memory: the tile leaves as N_TILE/64 bulk tensor stores issued by one thread
        // (per-thread NHWC stores put 32 scattered 16-byte pieces into every store instruction)
        const bool issuer = (warp == 4) && (lane == 0);
        // the bulk stores that last read this staging buffer (kPlainBufs tiles ago) have finished reading it
        if (issuer) tma_store_wait_read_n<Cfg::kPlainBufs - 1>();
        named_bar_sync(1, 32 * kEpiWarps);
        const uint32_t so = smem_u32(stage_out) + (it % Cfg::kPlainBufs) * Cfg::kPlainBufBytes;
#pragma unroll 1
        for (int cc = half; cc < N_TILE / 16; cc += kChunkStep) {
          uint32_t v[16];
          tmem_ld16(t_acc + cc * 16, v);
          tmem_ld_wait();
          if (cc + kChunkStep >= N_TILE / 16) release();
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            const int nl = cc * 16 + hlf * 8;                  // column inside the tile
            const int n0 = n_tile * N_TILE + nl;
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[hlf * 8 + e]);
            if (p.plain_bias && n0 < p.n_total) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.plain_bias + n0));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.plain_bias + n0) + 1);
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
              f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            if (p.plain_relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], p.plain_slope * f[e]);
            }
            // box nl/64 = [128 px][64 ch] bf16, 128-byte rows, SWIZZLE_128B: chunk ^= row & 7
            const uint32_t dst = so + (nl >> 6) * 16384 + row * 128 + ((((nl >> 3) & 7) ^ (row & 7)) << 4);
            st_shared_v4(dst, pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                         pack_bf16x2(f[6], f[7]));
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 32 * kEpiWarps);
        if (issuer) {
#pragma unroll
          for (int j = 0; j < N_TILE / 64; ++j) {
            const int col0 = n_tile * N_TILE + 64 * j;
            if (col0 >= p.n_total) break;
            // pixels outside the image / batch and channels past the tensor are clipped by the TMA unit
            if (col0 < p.Cin) {
              if (p.out0) tma_store_4d(&tmap_o0, so + j * 16384, col0, x0, y0, b);
            } else {
              if (p.out1) tma_store_4d(&tmap_o1, so + j * 16384, col0 - p.Cin, x0, y0, b);
            }
          }
          tma_store_commit();
        }
      } else {  // EPI_PLAIN: D -> bf16, columns [0,Cin) -> out0, [Cin, n_total) -> out1
#pragma unroll 1
        for (int cc = half; cc < N_TILE / 16; cc += kChunkStep) {
          uint32_t v[16];
          tmem_ld16(t_acc + cc * 16, v);
          tmem_ld_wait();
          if (cc + kChunkStep >= N_TILE / 16) release();
          if (valid) {
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              const int n0 = n_tile * N_TILE + cc * 16 + hlf * 8;
              if (n0 < p.n_total) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[hlf * 8 + e]);
                if (p.plain_bias) {
                  const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.plain_bias + n0));
                  const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.plain_bias + n0) + 1);
                  f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                  f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                }
                if (p.plain_relu) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], p.plain_slope * f[e]);
                }
                if (p.plain_f32) {
                  float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out0) + pix * p.n_total + n0);
                  dst[0] = make_float4(f[0], f[1], f[2], f[3]);
                  dst[1] = make_float4(f[4], f[5], f[6], f[7]);
                  continue;
                }
                uint4 o;
                o.x = pack_bf16x2(f[0], f[1]);
                o.y = pack_bf16x2(f[2], f[3]);
                o.z = pack_bf16x2(f[4], f[5]);
                o.w = pack_bf16x2(f[6], f[7]);
                if (p.plain_shuffle) {
                  // PixelShuffle(2) fused into the store (generator.py:24-26): 8 consecutive packed columns are 8
                  // channels of ONE sub-pixel
                  const int cps = p.n_total >> 2;
                  const int sub = n0 / cps, c = n0 - sub * cps;
                  const size_t opix = (static_cast<size_t>(b) * (2 * p.H) + (2 * y + (sub >> 1))) * (2 * p.W) +
                                      (2 * x + (sub & 1));
                  *reinterpret_cast<uint4*>(p.out0 + opix * cps + c) = o;
                } else if (p.o_map) {
                  // phase of a strided transposed conv: this tile's pixels are one parity class of dX
                  int f = b;
                  if (p.nd5) {
                    const int smp = b / p.T_out;
                    f = smp * p.o_T + (b - smp * p.T_out) * p.o_st + p.o_pt;
                  }
                  const size_t opix = (static_cast<size_t>(f) * p.o_H + (y * p.o_s + p.o_py)) * p.o_W + (x * p.o_s + p.o_px);
                  *reinterpret_cast<uint4*>(p.out0 + opix * p.Cin + n0) = o;
                } else if (n0 < p.Cin) {
                  if (p.out0) *reinterpret_cast<uint4*>(p.out0 + pix * p.Cin + n0) = o;
                } else {
                  if (p.out1) *reinterpret_cast<uint4*>(p.out1 + pix * (p.n_total - p.Cin) + (n0 - p.Cin)) = o;
                }
              }
            }
          }
        }
      }
      if (!released && !saved) release();   // warps without a chunk for this tile shape
    }
    if ((Cfg::kTmaStore && EPI != EPI_LSTM_BWD_GATES) || (EPI == EPI_PLAIN && p.plain_tma)) {
      if (warp == 4 && lane == 0) tma_store_wait_all();   // bulk stores complete before the CTA retires
    }
    if ((kProfEnabled && p.prof) && warp == 4 && lane == 0) {
      unsigned long long* q = p.prof + blockIdx.x * 16;
      q[4] = pa_wait; q[5] = pa_busy; q[6] = pa_barA; q[7] = pa_ld; q[8] = pa_math; q[9] = pa_barB;
      q[10] = pa_pre; q[11] = pa_iss; q[12] = pa_top;
    }
  }

  tc_fence_before();
  // pair: neither CTA may exit (or free TMEM) while the peer's MMAs / barrier arrives can still touch it
  if constexpr (kCta == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kCta>(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace plc
