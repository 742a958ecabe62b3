// MCB_TRMM.exe — the reference's TRMM.exe (TRMM.cpp): eigen-pairs of the Transition Rate Matrix a TRMM run left in
// output.h5, forward and adjoint, written to output_TRMM.h5 in the same directory.
//   usage: MCB_TRMM.exe <dir>/output.h5
#include <cstdio>
#include <string>

#include "trmm_eigen.h"

int main(int argc, char* argv[])
{
    if (argc != 2) { std::fprintf(stderr, "usage: %s <dir>/output.h5\n", argv[0]); return 2; }
    std::string error;
    if (!mcbhost::trmm_postprocess(argv[1], error)) { std::fprintf(stderr, "[ERROR] %s\n", error.c_str()); return 1; }
    return 0;
}
