/* tc_gemm.hpp -- interface of the tcgen05 split-precision spin GEMM (energy_tc.cu); fp32 solvers only. */
#pragma once
#include "device.hpp"

namespace sqb {

struct TcOperand { /* A (rows x K fp32) split into bf16 hi/mid/lo planes, [3][rowsPad][Kp], plus its TMA descriptor */
    TcOperand() : rows(0), rowsPad(0), K(0), Kp(0), ready(false) {}
    DevBuf<unsigned short> data;
    int rows, rowsPad, K, Kp;
    alignas(64) unsigned char map[128]; /* CUtensorMap */
    bool ready;
};
struct TcWorkspace {
    TcWorkspace() : dev(NULL) {}
    DevBuf<unsigned short> qbf; /* Q widened to bf16, [roundUp(m,128)][Kp] */
    DevBuf<float> cbuf;         /* C for the energy path */
    const B200Device *dev;
};

bool tcEnabled(); /* false when SQAOD_B200_NO_TC is set (A/B testing) or the driver lacks cuTensorMapEncodeTiled */
void tcPrepareOperand(const B200Device &dev, TcOperand &op, const float *d_A, int ldA, int rows, int K);
/* C[y][i] = sum_k Q[y][k] A[i][k],  y < m, i < A.rows */
void tcSpinGemm(const B200Device &dev, float *d_C, int ldc, const TcOperand &A, const signed char *d_Q, int ldq, int m, TcWorkspace &ws);
/* the same product from spins the caller keeps in bf16 ([roundUp(m,128)][A.Kp], padding zero), e.g. maintained by a flip kernel */
void tcSpinGemmBf16(const B200Device &dev, float *d_C, int ldc, const TcOperand &A, const unsigned short *d_Qbf, int m);
/* (re)build such a bf16 copy from int8 spins; allocates / zero-fills `qbf` when it is too small */
void tcWidenSpins(const B200Device &dev, DevBuf<unsigned short> &qbf, const TcOperand &A, const signed char *d_Q, int ldq, int m);
/* E_b = alpha (sum_i v_bi (g_i + sum_j A_ij u_bj) + f.u_b) + beta0 */
void tcBatchedEnergy(const B200Device &dev, float *d_E, const TcOperand &A, const signed char *d_u, int ldu, const signed char *d_v, int ldv,
                     const float *d_g, const float *d_f, int nBatch, float alpha, float beta0, TcWorkspace &ws);

} // namespace sqb
