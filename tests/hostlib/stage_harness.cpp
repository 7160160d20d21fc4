// Host build of the plan compiler + per-proof stages (tests only; never shipped): runs ONE proof on
// the CPU through exactly the code the kernels run, so the logic can be checked without a GPU.
#include <string.h>
#include <vector>
#include "plan_build.h"
#include "stages.cuh"
#include "tower.cuh"
using namespace h2v;
static std::vector<u8> g_blob; static PlanInfo g_info; static std::string g_err;
extern "C" {
int s_build(const u8* params, size_t pl, int pf, const u8* vk, size_t vl, int vf, int mo, int hash) {
  g_info.lint.clear();
  return build_plan(params, pl, pf, vk, vl, vf, mo, hash, g_blob, g_info, g_err);
}
int s_build_m(const u8* params, size_t pl, int pf, const u8* vk, size_t vl, int vf, int mo, int hash, u32 circuit_instances) {
  g_info.lint.clear();
  return build_plan(params, pl, pf, vk, vl, vf, mo, hash, g_blob, g_info, g_err, circuit_instances);
}
const char* s_err() { return g_err.c_str(); }
// VK lint findings of the last successful s_build (PlanInfo::lint), newline-separated
static std::string g_lint;
const char* s_lint() { g_lint.clear(); for (auto& f : g_info.lint) g_lint += f + "\n"; return g_lint.c_str(); }
void s_info(u32* out) { memcpy(out, &g_info, 8 * sizeof(u32)); }
// returns status; outputs canonical LE: challenges [C][32], right [P][32], shared [Sh][32], left [n_mo][32],
// L,R affine canonical x|y (zeros = identity) and verdict of the pairing in *pair_ok
int s_verify_one(const u8* proof, u32 len, const u8* inst, u32 inst_total, const u32* col_len, int ncols,
                 u8* challenges, u8* right, u8* shared, u8* left, u8* LR, int* pair_ok) {
  PlanView pv{g_blob.data()}; const PlanHeader& hd = pv.h();
  u32 st = ST_OK;
  if (ncols >= 0 && (u32)ncols != hd.n_inst_cols) return ST_INVALID_INSTANCES;
  if (!col_len && hd.n_inst_cols && inst_total % hd.n_inst_cols) return ST_INVALID_INSTANCES;
  if (!hd.n_inst_cols && inst_total) return ST_INVALID_INSTANCES;
  std::vector<G1Affine> pts(hd.n_points);
  u32 bad = H2V_NO_BAD_ITEM;
  const u32* pt_item = pv.sec<u32>(hd.off_pt_item);
  for (u32 s = 0; s < hd.n_points; s++) {
    if (!decompress_stage(pv, proof, len, s, pts[s])) { pts[s].x = Fq::zero(); pts[s].y = Fq::zero(); if (pt_item[s] < bad) bad = pt_item[s]; }
  }
  std::vector<Fr> vals(hd.n_vals);
  bool inst_bad = false;
  if (hd.hash == HASH_BLAKE2B) bad = transcript_stage<Blake2b>(pv, proof, len, inst, inst_total, pts.data(), vals.data(), 0, 1, bad, inst_bad);
  else bad = transcript_stage<Keccak256>(pv, proof, len, inst, inst_total, pts.data(), vals.data(), 0, 1, bad, inst_bad);
  for (u32 c = 0; c < hd.n_challenges; c++) vals[hd.v_chal + c].to_canonical().store_le(challenges + 32 * c);
  if (inst_bad) return ST_INVALID_INSTANCES;
  if (bad != H2V_NO_BAD_ITEM) return bad < hd.first_mo_item ? ST_TRANSCRIPT : ST_OPENING;
  u32 max_len = 0; if (col_len) for (u32 c = 0; c < hd.n_inst_cols; c++) max_len = std::max(max_len, col_len[c]); else if (hd.n_inst_cols) max_len = inst_total / hd.n_inst_cols;
  std::vector<Fr> scratch(hd.inst_max_rot + max_len + hd.inst_min_rot_abs + 1), R_(hd.n_points), S_(hd.n_shared), L_(hd.n_mo);
  ScalarIO io{0, 1, vals.data(), scratch.data(), R_.data(), S_.data(), L_.data()};
  st = scalar_stage(pv, io, inst, col_len, inst_total);
  if (st != ST_OK) return st;
  for (u32 i = 0; i < hd.n_points; i++) R_[i].to_canonical().store_le(right + 32 * i);
  for (u32 i = 0; i < hd.n_shared; i++) S_[i].to_canonical().store_le(shared + 32 * i);
  for (u32 i = 0; i < hd.n_mo; i++) L_[i].to_canonical().store_le(left + 32 * i);
  // per-proof accumulators by plain double-and-add
  G1Jac Lacc = G1Jac::identity(), Racc = G1Jac::identity();
  const G1Affine* sh = pv.sec<G1Affine>(hd.off_shared_pts);
  for (u32 i = 0; i < hd.n_points; i++) { Fr k = R_[i].to_canonical(); Racc = g1_add(Racc, g1_mul_canonical(pts[i], k.l)); }
  for (u32 i = 0; i < hd.n_shared; i++) { if (sh[i].x.is_zero() && sh[i].y.is_zero()) continue; Fr k = S_[i].to_canonical(); Racc = g1_add(Racc, g1_mul_canonical(sh[i], k.l)); }
  for (u32 i = 0; i < hd.n_mo; i++) { Fr k = L_[i].to_canonical(); Lacc = g1_add(Lacc, g1_mul_canonical(pts[hd.n_points - hd.n_mo + i], k.l)); }
  G1Affine La, Ra; g1_to_affine(Lacc, La); g1_to_affine(Racc, Ra);
  La.x.to_canonical().store_le(LR); La.y.to_canonical().store_le(LR + 32); Ra.x.to_canonical().store_le(LR + 64); Ra.y.to_canonical().store_le(LR + 96);
  bool ok = pairing_check2(Lacc, Racc, pv.sec<G2Line>(hd.off_lines0), pv.sec<G2Line>(hd.off_lines1));
  *pair_ok = ok;
  return ok ? ST_OK : ST_CONSTRAINT_SYSTEM_FAILURE;
}
}
