/* sqaodc/sqaodc.h -- drop-in for the reference's umbrella header (sqaodc/sqaodc.h:1-66): lets code written against
 * libsqaodc_cuda -- in particular the reference's CPython glue, sqaodc/pyglue/{pyglue.h,annealer.inc,bf_searcher.inc,
 * formulas.inc} and sqaodpy/sqaod/cuda/src/cuda_*.cpp -- compile UNMODIFIED against libsqaod_b200.so:
 *     g++ -I<repo>/include -I<reference root> ... sqaodpy/sqaod/cuda/src/cuda_dg_annealer.cpp -lsqaod_b200
 * (recipe: oracle/Makefile target `glue`; exercised by tests/test_reference_suite_gpu.py).
 * The error macros carry the reference's names (sqaodc/common/defines.h:43-47). */
#pragma once
#include <sqaod_b200/sqaod_api.hpp>
#include <assert.h>

#ifndef throwError
#define abort_(...) ::sqaod::abortAt(__FILE__, __LINE__, __VA_ARGS__)
#define abortIf(cond, ...) if (cond) ::sqaod::abortAt(__FILE__, __LINE__, __VA_ARGS__)
#define throwError(...) ::sqaod::throwErrorAt(__FILE__, __LINE__, __VA_ARGS__)
#define throwErrorIf(cond, ...) if (cond) ::sqaod::throwErrorAt(__FILE__, __LINE__, __VA_ARGS__)
#endif
