// k0_host.cc — tables of the K0 core and its CPU execution (test scaffold: the product runs K0 on the GPU).
#include "k0_host.h"
#include <cstring>
#include "hevc_cabac.h"
#include "hevc_scan.h"

namespace hc {

namespace {
const uint8_t kSigCtx4x4_[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};
const uint8_t kMode422_[35] = {0,  1,  2,  2,  2,  2,  3,  5,  7,  8,  10, 12, 13, 15, 17, 18, 19, 20,
                               21, 22, 23, 23, 24, 24, 25, 25, 26, 27, 27, 28, 28, 29, 29, 30, 31};
const int8_t kQpc420_[14] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37};

struct TablesBuilder {
  k0::Tables t;
  TablesBuilder() {
    memset(&t, 0, sizeof(t));
    for (int i = 0; i < 256; i++) t.recip[i] = (uint32_t)((1ull << 34) / (uint64_t)(256 + i) + 1);
    memcpy(t.range_lps, detail::kRangeLps, sizeof(t.range_lps));
    memcpy(t.next_state, detail::kTransitions.next, sizeof(t.next_state));
    uint8_t init[CTX_COUNT];
    ctx_init_values(init);
    static_assert((int)k0::CX_COUNT <= (int)CTX_COUNT, "context layout");
    static_assert((int)k0::CX_G2 == (int)CTX_G2 && (int)k0::CX_SIG == (int)CTX_SIG && (int)k0::CX_LAST_Y == (int)CTX_LAST_Y &&
                  (int)k0::CX_TSKIP == (int)CTX_TSKIP && (int)k0::CX_SPLIT_TRANSFORM == (int)CTX_SPLIT_TRANSFORM, "context layout");
    memcpy(t.ctx_init, init, k0::CX_COUNT);
    const ScanTables& st = scan_tables();
    for (int s = 0; s < 3; s++)
      for (int k = 0; k < 16; k++) {
        t.scan_pos[s][k] = (uint8_t)(st.order[2][s][k].x | (st.order[2][s][k].y << 2));
        t.inv_pos[s][k] = st.inverse[2][s][k];
      }
    for (int l = 0; l <= 3; l++)
      for (int s = 0; s < 3; s++)
        for (int i = 0; i < (1 << (2 * l)); i++) {
          t.scan_sub[l][s][i] = (uint8_t)(st.order[l][s][i].x | (st.order[l][s][i].y << 3));
          t.inv_sub[l][s][i] = st.inverse[l][s][i];
        }
    // significance contexts: same construction as SigCtxTables in hevc_parse.cc (9.3.4.2.5)
    for (int c = 0; c < 2; c++)
      for (int s = 0; s < 3; s++)
        for (int k = 0; k < 16; k++) {
          const int xP = st.order[2][s][k].x, yP = st.order[2][s][k].y;
          t.sig_b4[c][s][k] = (uint8_t)((c ? 27 : 0) + kSigCtx4x4_[(yP << 2) + xP]);
          for (int cls = 0; cls < 3; cls++)
            for (int nz = 0; nz < 2; nz++)
              for (int prev = 0; prev < 4; prev++) {
                int v;
                switch (prev) {
                  case 0: v = (xP + yP >= 3) ? 0 : (xP + yP > 0) ? 1 : 2; break;
                  case 1: v = (yP == 0) ? 2 : (yP == 1) ? 1 : 0; break;
                  case 2: v = (xP == 0) ? 2 : (xP == 1) ? 1 : 0; break;
                  default: v = 2; break;
                }
                if (c == 0) {
                  if (nz) v += 3;
                  v += cls == 0 ? 9 : (cls == 1 ? 15 : 21);
                } else {
                  v += cls == 2 ? 12 : 9;
                }
                t.sig_sb[c][cls][nz][prev][s][k] = (uint8_t)((c ? 27 : 0) + v);
              }
        }
    memcpy(t.mode422, kMode422_, 35);
    memcpy(t.qpc420, kQpc420_, 14);
  }
};
}  // namespace

const k0::Tables& k0_tables() {
  static const TablesBuilder b;
  return b.t;
}

std::string k0_prepare(const uint8_t* data, size_t size, int stream_format, K0HostPicture& out) {
  // one parser per thread, reset between items: building its parameter-set tables (~700 KB) was most of the 0.13 ms
  // this function cost per 512 x 512 item — the host work every device-parsed item still needs
  thread_local HevcIntraParser parser;
  parser.reset();
  parser.set_collect_only(true);
  std::string e;
  if (stream_format == 0) e = parser.push_length_prefixed(data, size);
  else if (stream_format == 1) e = parser.push_annexb(data, size);
  else e = parser.push_nal(data, size);
  if (!e.empty()) return e;
  return parser.take_k0(out);
}

std::unique_ptr<PictureRecords> k0_parse_on_cpu(const K0HostPicture& hp, std::string* err) {
  if (!hp.eligible) { if (err) *err = "picture is not eligible for K0: " + hp.why_not; return nullptr; }
  k0::Pic p = hp.pic;
  const size_t nctb = (size_t)p.ctbs_w * p.ctbs_h;
  std::vector<uint8_t> ct_depth((size_t)p.w8 * p.h8, 0), ipm((size_t)p.w4 * p.h4, 1), ipm_c((size_t)p.w4 * p.h4, 1);
  std::vector<uint8_t> wpp((size_t)p.ctbs_h * k0::CTX_BYTES, 0), edge((size_t)p.w4 * p.h4, 0);
  std::vector<int8_t> qp((size_t)p.w8 * p.h8, 0);
  std::vector<int> progress(p.ctbs_h, 0);
  std::vector<hc_ctu> ctus(nctb);
  std::vector<hc_blk> blks(nctb * p.blk_cap_ctb);
  std::vector<hc_tb> tbs(nctb * p.tb_cap_ctb);
  std::vector<hc_coeff> coeffs(nctb * p.coeff_cap_ctb);
  std::vector<uint32_t> lists[4];
  for (auto& l : lists) l.resize(tbs.size());
  unsigned int counts[4] = {0, 0, 0, 0};
  int error = 0;
  memset(ctus.data(), 0, ctus.size() * sizeof(hc_ctu));
  p.bytes = hp.bytes.data(); p.slices = hp.slices.data(); p.ctb_slice = hp.ctb_slice.data(); p.ctu_static = hp.ctu_static.data();
  p.ct_depth = ct_depth.data(); p.ipm = ipm.data(); p.ipm_c = ipm_c.data(); p.wpp_ctx = wpp.data(); p.progress = progress.data();
  p.error = &error; p.qp_map = qp.data(); p.edge_map = edge.data(); p.ctus = ctus.data(); p.blks = blks.data(); p.tbs = tbs.data();
  p.coeffs = coeffs.data();
  for (int l = 0; l < 4; l++) p.tb_lists[l] = lists[l].data();
  p.tb_counts = counts;
  p.pic_index = 0; p.tb_global_base = 0;

  std::unique_ptr<k0::Scratch> scratch(new k0::Scratch);
  memset(scratch.get(), 0, sizeof(k0::Scratch));
  for (const k0::Chain& ch : hp.chains) {      // row order satisfies every wavefront dependency
    k0::Parser ps;
    memset(&ps, 0, sizeof(ps));
    ps.T = &k0_tables();
    ps.S = scratch.get();
    ps.run_chain(&p, hp.subs.data(), ch.first_sub, ch.nsubs);
    if (error) break;
  }
  if (error) { if (err) *err = error == k0::ERR_CAPACITY ? "K0: per-CTB capacity exceeded" : "K0: malformed slice data"; return nullptr; }
  for (size_t c = 0; c < nctb; c++) k0::finish_ctb(p, (int)c, 0, 1);   // second half of K0 (k0_finish_kernel on the device)

  // compact the fixed-capacity CTB slices into the host parser's sequential record form
  std::unique_ptr<PictureRecords> rec(new PictureRecords);
  rec->pic = hp.hpic;
  rec->scaling = hp.scaling;
  rec->ctus = ctus;
  rec->edge_map = edge;
  rec->qp_map.assign(qp.begin(), qp.end());
  for (size_t c = 0; c < nctb; c++) {
    // transform blocks of this CTB in parse order, with their coefficients and residual offsets renumbered
    std::vector<uint32_t> new_resid(p.tb_cap_ctb, 0);
    uint32_t ntb = 0;
    // the number of TBs used in this CTB is not stored: walk while residual offsets stay inside the CTB slice
    // and were written (K0 fills slices front to back); use the blocks to know which TBs exist
    uint32_t used_tb = 0;
    for (int comp = 0; comp < 3; comp++)
      for (uint32_t k = 0; k < ctus[c].blk_count[comp]; k++)
        if (blks[ctus[c].blk_first[comp] + k].flags & HC_BLK_HAS_RESID) used_tb++;
    for (uint32_t t = 0; t < used_tb; t++) {
      hc_tb tb = tbs[c * p.tb_cap_ctb + t];
      const uint32_t old_resid = tb.resid_off;
      const hc_coeff* src = coeffs.data() + tb.coeff_off;
      tb.coeff_off = (uint32_t)rec->coeffs.size();
      tb.resid_off = (uint32_t)rec->resid_count;
      tb.pic = 0;
      rec->coeffs.insert(rec->coeffs.end(), src, src + tb.ncoeff);
      rec->resid_count += (uint64_t)1 << (2 * tb.log2);
      rec->tbs_by_size[tb.log2 - 2]++;
      // remember the mapping for the blocks
      new_resid[t] = tb.resid_off;
      (void)old_resid;
      rec->tbs.push_back(tb);
      ntb++;
    }
    // blocks: Y, Cb, Cr lists concatenated; residual offsets mapped through the TB that owns them
    for (int comp = 0; comp < 3; comp++) {
      const uint32_t first = ctus[c].blk_first[comp];
      rec->ctus[c].blk_first[comp] = (uint32_t)rec->blks.size();
      for (uint32_t k = 0; k < ctus[c].blk_count[comp]; k++) {
        hc_blk b = blks[first + k];
        if (b.flags & HC_BLK_HAS_RESID) {
          uint32_t t = 0;
          for (; t < used_tb; t++)
            if (tbs[c * p.tb_cap_ctb + t].resid_off == b.resid_off) break;
          b.resid_off = new_resid[t < used_tb ? t : 0];
        }
        rec->blks.push_back(b);
      }
    }
  }
  rec->pic.blk_count = (uint32_t)rec->blks.size();
  rec->pic.tb_count = (uint32_t)rec->tbs.size();
  rec->pic.coeff_count = (uint32_t)rec->coeffs.size();
  rec->pic.resid_count = rec->resid_count;
  if (err) err->clear();
  return rec;
}

}  // namespace hc
