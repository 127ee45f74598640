// Instantiated (nx, nu, threads per (problem, segment)) triples -- the single source of truth for _build.py (one
// object file each) and for the registry in pdplqr.cu.  The BASELINE.json configs use (12,4), (4,1) and (30,10); the
// rest cover generic sizes: any other (nx, nu) is zero / identity padded to the cheapest triple that contains it
// (pdplqr_create), so the list also bounds the largest supported problem.
#pragma once
#define PDPLQR_INST_LIST(X) \
    X(12, 4, 32) X(4, 1, 32) X(30, 10, 128) X(2, 1, 32) X(3, 2, 32) X(6, 3, 32) X(8, 8, 32) X(6, 2, 32) X(8, 4, 32) \
    X(16, 4, 32)
