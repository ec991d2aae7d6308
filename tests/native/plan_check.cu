// Host-only check of the launch-plan arithmetic of the fused / split-K kernels (common.cuh): every EfficientNet-B0 block at
// 256x256 and 512x512 either gets a plan that respects the sm_100a budgets (227 KB shared memory per CTA, 512 TMEM columns,
// portable cluster size, one wave of CTAs) or is rejected so that the engine keeps the generic launches.
// Built and run by tests/test_plans.py (nvcc, no GPU needed).
#include <cstdio>
#include <cstring>

#include "common.cuh"

using namespace hp;

struct Blk { int k, s, e, cin, cout; };
static const Blk kBlocks[16] = {{3, 1, 1, 32, 16},  {3, 2, 6, 16, 24},   {3, 1, 6, 24, 24},   {5, 2, 6, 24, 40},
                                {5, 1, 6, 40, 40},  {3, 2, 6, 40, 80},   {3, 1, 6, 80, 80},   {3, 1, 6, 80, 80},
                                {5, 1, 6, 80, 112}, {5, 1, 6, 112, 112}, {5, 1, 6, 112, 112}, {5, 2, 6, 112, 192},
                                {5, 1, 6, 192, 192}, {5, 1, 6, 192, 192}, {5, 1, 6, 192, 192}, {3, 1, 6, 192, 320}};

static int fail(const char* what, int size, int i) {
  std::printf("FAIL %s (image %d, block %d)\n", what, size, i);
  return 1;
}

int main() {
  int bad = 0, n_mb = 0, n_ed = 0, n_pk = 0;
  for (int size : {256, 512}) {
    int H = size / 2;
    for (int i = 0; i < 16; ++i) {
      const Blk& b = kBlocks[i];
      const int Ho = (H + b.s - 1) / b.s, cexp = b.cin * b.e;
      if (b.e != 1) {
        MbSpec m;
        std::memset(&m, 0, sizeof(m));
        m.H = m.W = H; m.Ho = m.Wo = Ho; m.cin = b.cin; m.cexp = cexp; m.cout = b.cout; m.cse = b.cin / 4;
        m.k = b.k; m.stride = b.s; m.skip = (b.s == 1 && b.cin == b.cout);
        if (mb_plan(m)) {
          ++n_mb;
          if (m.smem_bytes > 227 * 1024) bad += fail("mbconv smem", size, i);
          if (m.tmem_cols > 512 || m.d2_col0 + m.MTo * m.d2_pitch > m.tmem_cols) bad += fail("mbconv tmem", size, i);
          if (m.cl < 1 || m.cl > MB_CL_MAX || m.nmine > MB_MAX_MINE || m.cl * m.nmine < m.nsl) bad += fail("mbconv cluster", size, i);
          if (m.rows_own * m.cl < m.Po) bad += fail("mbconv rows", size, i);
          if (16 * MB_STAGE_WARP_BYTES > m.off_dw) bad += fail("mbconv staging", size, i);
          if (mb_part_bytes(m, 16) != (size_t)16 * m.cl * m.Po * m.cout * 4) bad += fail("mbconv scratch", size, i);
        } else if (size == 256 && i >= 6) {
          bad += fail("mbconv plan missing for a small-map block", size, i);
        }
        EdSpec e;
        std::memset(&e, 0, sizeof(e));
        e.B = 16; e.H = e.W = H; e.Ho = e.Wo = Ho; e.cin = b.cin; e.cexp = cexp; e.k = b.k; e.stride = b.s;
        if (ed_plan(e)) {
          ++n_ed;
          if (e.smem_bytes > 227 * 1024) bad += fail("expdw smem", size, i);
          if (e.cexp > 256 || e.cin > 64) bad += fail("expdw operand limits", size, i);
          if ((e.TO - 1) * b.s + b.k > ED_WIN) bad += fail("expdw window", size, i);
          if (e.tiles_x * e.TO < e.Wo || e.tiles_y * e.TO < e.Ho || e.total_tiles != 16 * e.tiles_per_img) bad += fail("expdw tiling", size, i);
          if (e.tiles_per_img != ed_tiles_per_img(b.k, b.s, Ho, Ho)) bad += fail("expdw tile count", size, i);
          if ((e.e_pitch / 16) % 2 == 0) bad += fail("expdw pitch parity", size, i);
        } else if (i >= 1 && i <= 5) {
          bad += fail("expdw plan missing for a large-map block", size, i);
        }
      }
      for (int batch : {1, 2, 4, 16}) {
        PkSpec p;
        std::memset(&p, 0, sizeof(p));
        p.M = batch * Ho * Ho; p.N = b.cout; p.K = cexp; p.rows_per_img = Ho * Ho;
        if (pk_plan(p, 148)) {
          ++n_pk;
          if (p.smem_bytes > 226 * 1024 || p.tmem_cols > 512) bad += fail("projk budgets", size, i);
          if (p.S < 2 || p.S > 6 || p.nmine > PK_MAX_MINE || p.S * p.nmine < p.nkb) bad += fail("projk cluster", size, i);
          if (p.m_tiles * p.S > 148) bad += fail("projk more than one wave", size, i);
          if (p.rows_own * p.S < 128) bad += fail("projk rows", size, i);
        } else if (size == 256 && i >= 6 && batch <= 4) {
          bad += fail("projk plan missing for a latency plan", size, i);
        }
      }
      H = Ho;
    }
  }
  std::printf("plans: mbconv %d, expdw %d, projk %d, failures %d\n", n_mb, n_ed, n_pk, bad);
  return bad ? 1 : 0;
}
