/*
 * oracle/ref_pack.h — TEST INFRASTRUCTURE ONLY.
 * The conversions between the reference's SBR / PS structs and the flat records of the C-ABI are part of the reference-side
 * drop-in (libxaac_b200/dropin/ixheaacd_b200_pack.h, XAAC_* offsets of include/xaac_b200.h); the taps and the shim that drives
 * the compiled reference from records use the same code.  tests/test_abi.py checks that the XO_* offsets of
 * oracle/src/xaac_oracle.h and the XAAC_* offsets agree.
 */
#ifndef XAAC_REF_PACK_H
#define XAAC_REF_PACK_H
#include "ref_headers.h"
#include "../libxaac_b200/dropin/ixheaacd_b200_pack.h"
#include "src/xaac_oracle.h"
#endif
