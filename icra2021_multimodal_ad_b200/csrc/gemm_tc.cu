// tcgen05 GEMM -- placeholder until the tensor-core engine lands (next commit).
#include "mmad_internal.cuh"
namespace mmad {
int tc_available() { return 0; }
int gemm_tc_tile_n() { return 128; }
int tc_make_operand_map(CUtensorMap*, const __half*, int, int, int, int) { set_error("tcgen05 path not built"); return MMAD_E_UNSUPPORTED; }
int gemm_tc(const TcOperand&, const TcOperand&, int, int, int, int, const Epilogue&, cudaStream_t) { set_error("tcgen05 path not built"); return MMAD_E_UNSUPPORTED; }
}
