// lhs_plan.h -- host-side plan of FSILS_LHS_CREATE (L/LHS.f:113-288)
#pragma once
#include <stdint.h>
#include <vector>

namespace svfsi {

struct LhsPlan {
  struct Nbr {
    int iP;                 // 0-based neighbour rank
    std::vector<int> ptr;   // 0-based reordered local node ids, pair-consistent order
  };
  std::vector<int> map;     // svFSI local id (0-based) -> reordered id (0-based)
  int mynNo = 0, shnNo = 0;
  std::vector<Nbr> nbr;     // ascending iP
};

// aNodes = [nranks][maxnNo] zero-padded 1-based global ids (the MPI_ALLGATHERV
// buffer of L/LHS.f:125).
LhsPlan lhs_plan(int rank, int nranks, int gnNo, int nNo, int maxnNo, const int32_t *aNodes);

}  // namespace svfsi
