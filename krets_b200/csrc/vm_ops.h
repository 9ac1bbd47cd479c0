// Lane-program ("lane VM") instruction set shared by the host plan compiler and the sm_100a interpreter kernel.
//
// The host turns one circuit + analysis mode into a branch-free straight-line program over
//   scr[slot]      per-lane on-chip scratch cells (shared memory column of the lane's thread)
//   B[idx]         lane-invariant constants (element-derived coefficients, source tables) — broadcast loads
//   IN[idx][lane]  per-lane inputs (sweep value, frequency, Monte-Carlo element values) — coalesced loads
//   G[idx][lane]   per-lane HBM value planes (spilled L/U/y values) — coalesced loads/stores
//   X[row][u][lane] result rows (also the NR guess / previous time point)
// Every lane executes the identical op sequence, so a warp never diverges and every G/IN/X access is one
// contiguous 256 B (f64) or 512 B (c64) segment per warp.
//
// Encoding: 2 x uint32 per op:  w0 = opc | d<<8 | a<<16 | b<<24 ;  w1 = imm
#pragma once
#include <stdint.h>

enum KrbOp : uint32_t {
    KOP_END = 0,
    KOP_LDB = 1,    // scr[d] = B[imm]
    KOP_LDBS = 2,   // scr[d] = B[imm + row*b_row_stride + min(iter,1)*b_iter_stride]   (per-time-step tables)
    KOP_LDI = 3,    // scr[d] = IN[imm][lane]              (real plane; complex programs get (v, 0))
    KOP_LDG = 4,    // scr[d] = G[imm][lane]
    KOP_STG = 5,    // G[imm][lane] = scr[a]
    KOP_LDX = 6,    // scr[d] = X[b&1 ? guess_row : prev_row][imm][lane]; if the guess map is empty: (b&2) ? 0.5 : 0.0
    KOP_XOUT = 7,   // deliver unknown imm: NR convergence test against the guess (b&1: current tolerance), store to X[cur_row]
    KOP_ZERO = 8,   // scr[d] = 0
    KOP_MOV = 9,    // scr[d] = scr[a]
    KOP_NEG = 10,   // scr[d] = -scr[a]
    KOP_ADD = 11,   // scr[d] = scr[a] + scr[b]
    KOP_SUB = 12,   // scr[d] = scr[a] - scr[b]
    KOP_MUL = 13,   // scr[d] = scr[a] * scr[b]
    KOP_FNMA = 14,  // scr[d] -= scr[a] * scr[b]
    KOP_RCP = 15,   // scr[d] = 1 / scr[a]   (static pivot: zero / non-finite sets KRB_LANE_ZERO_PIVOT)
    KOP_INV = 16,   // scr[d] = 1.0 / re(scr[a])           (stamp: g = 1.0 / value, stampable.rs:232)
    KOP_DIVR = 17,  // scr[d] = re(scr[a]) / re(scr[b])    (stamp: value / h, stampable.rs:449)
    KOP_MULJ = 18,  // scr[d] = (0, re(scr[a]))            (stamp: admittance (0, 2*pi*f*C), stampable.rs:379-382)
    KOP_DIOG = 19,  // scr[d] = diode conductance at vd = scr[a]; B[imm..imm+3] = {is/(n*Vt), n*Vt, is, vcrit}  (diode.rs:42-48)
    KOP_DIOI = 20,  // scr[d] = diode equivalent current at vd = scr[a]                                          (diode.rs:50-61)
    KOP_LDK = 21,   // scr[d] = the f64 constant stored in the NEXT op slot (both words); that slot is skipped.  Never at chunk position 31.
    KOP_NOP = 22,   // padding
    KOP_WAIT = 23,  // wait for the asynchronous loads (LDI/LDG/LDX/LDBS are issued as cp.async straight into the scratch slot)
    KOP_MULG = 24,  // scr[d] = scr[a] * scr[b], a scaled factor entry u^ = u * (1/pivot): also tracks the lane's pivot growth (KRB_LANE_INACCURATE)
    // level-1 NMOS at v_gs = re(scr[a]), v_ds = re(scr[b]); B[imm..imm+2] = {Vth, beta, lambda}  (nmosfet.rs:51-102; extension, see KRB_ELEM_FLAG_LEVEL1)
    KOP_MOSGM = 25,   // scr[d] = g_m
    KOP_MOSGDS = 26,  // scr[d] = g_ds
    KOP_MOSID = 27,   // scr[d] = i_d
    KOP_COUNT = 28
};

static inline uint32_t krb_op_w0(uint32_t opc, uint32_t d, uint32_t a, uint32_t b) { return opc | (d << 8) | (a << 16) | (b << 24); }

#define KRB_MAX_SLOTS 256
