// lhs_plan.cpp -- host-side node reordering and halo schedule of FSILS_LHS_CREATE
// (reference: Code/Source/svFSILS/LHS.f:113-288).  Pure host code, no CUDA.
//
// Unlike the reference, which exchanges the per-pair node lists with
// MPI_SEND/MPI_RECV (L/LHS.f:254-287), every rank derives both directions
// locally from the two all-gathered node tables: for a pair (lo, hi) the
// common list is "hi's reordered node list filtered by membership in lo", which
// either side can evaluate once it knows every rank's reordered list -- and a
// rank can compute any other rank's reordered list from the first table alone.
#include "lhs_plan.h"

#include <algorithm>
#include <stdexcept>

namespace svfsi {

// L/LHS.f:134-175 for rank tF (0-based): returns that rank's reordered global-id
// list ("ltg" in the reference) and mynNo / shnNo.
static void reorder_rank(int tF, int nranks, int gnNo, int maxnNo, const int32_t *aNodes,
                         std::vector<int32_t> &gtl, std::vector<int32_t> &ltgNew,
                         int &mynNo, int &shnNo) {
  const int32_t *mineIn = aNodes + (size_t)tF * maxnNo;
  int nNo = 0;
  while (nNo < maxnNo && mineIn[nNo] != 0) nNo++;
  std::vector<int32_t> mine(mineIn, mineIn + nNo);
  // gtl: global id -> local position + 1 on rank tF (0 = not held)
  for (int a = 0; a < nNo; a++) gtl[mine[a] - 1] = a + 1;
  ltgNew.assign(nNo, 0);
  mynNo = nNo;
  shnNo = 0;
  for (int i = nranks - 1; i >= 0; i--) {
    if (i == tF) continue;
    const int32_t *other = aNodes + (size_t)i * maxnNo;
    for (int a = 0; a < maxnNo; a++) {
      int Ac = other[a];
      if (Ac == 0) break;
      int ai = gtl[Ac - 1];
      if (ai != 0 && mine[ai - 1] != 0) {
        if (i < tF) {
          ltgNew[shnNo++] = Ac;  // shared with a lower rank: front
        } else {
          ltgNew[--mynNo] = Ac;  // shared with a higher rank: back, filled downward
        }
        mine[ai - 1] = 0;
      }
    }
  }
  int j = shnNo;
  for (int a = 0; a < nNo; a++)
    if (mine[a] != 0) ltgNew[j++] = mine[a];
  if (j != mynNo) throw std::runtime_error("FSILS: Unexpected behavior in node reordering");
  for (int a = 0; a < nNo; a++) gtl[mineIn[a] - 1] = 0;  // leave gtl clean
  (void)gnNo;
}

LhsPlan lhs_plan(int rank, int nranks, int gnNo, int nNo, int maxnNo, const int32_t *aNodes) {
  LhsPlan p;
  p.map.resize(nNo);
  if (nranks == 1) {
    for (int a = 0; a < nNo; a++) p.map[a] = a;
    p.mynNo = nNo;
    p.shnNo = 0;
    return p;
  }
  std::vector<int32_t> gtl(gnNo, 0);
  std::vector<std::vector<int32_t>> ltgNew(nranks);
  std::vector<int> my(nranks), sh(nranks);
  for (int r = 0; r < nranks; r++)
    reorder_rank(r, nranks, gnNo, maxnNo, aNodes, gtl, ltgNew[r], my[r], sh[r]);
  p.mynNo = my[rank];
  p.shnNo = sh[rank];

  // map: svFSI local id -> reordered id (L/LHS.f:183-191)
  const std::vector<int32_t> &mineNew = ltgNew[rank];
  for (int a = 0; a < nNo; a++) gtl[mineNew[a] - 1] = a + 1;
  const int32_t *gNodes = aNodes + (size_t)rank * maxnNo;
  for (int a = 0; a < nNo; a++) p.map[a] = gtl[gNodes[a] - 1] - 1;

  // neighbours in ascending rank (L/LHS.f:219-249); list order = the HIGHER
  // rank's reordered numbering (L/LHS.f:251-288)
  std::vector<char> held;
  for (int i = 0; i < nranks; i++) {
    if (i == rank) continue;
    const std::vector<int32_t> &oth = ltgNew[i];
    LhsPlan::Nbr nb;
    nb.iP = i;
    if (i > rank) {
      for (int32_t Ac : oth) {
        int ai = gtl[Ac - 1];
        if (ai != 0) nb.ptr.push_back(ai - 1);
      }
    } else {
      if (held.empty()) held.assign(gnNo, 0);
      for (int32_t Ac : oth) held[Ac - 1] = 1;
      for (int32_t Ac : mineNew)
        if (held[Ac - 1]) nb.ptr.push_back(gtl[Ac - 1] - 1);
      for (int32_t Ac : oth) held[Ac - 1] = 0;
    }
    if (!nb.ptr.empty()) p.nbr.push_back(std::move(nb));
  }
  return p;
}

}  // namespace svfsi
