// codegen.hpp -- drop-in mirror of the reference's code-generation entry points
// (tinympc/TinyMPC/src/tinympc/codegen.hpp:9-19): same names, arguments and return values.
#pragma once
#include "types.hpp"

extern "C" {

// Writes <output_dir>/tinympc/tiny_data.hpp, <output_dir>/src/tiny_data.cpp and <output_dir>/src/tiny_main.cpp exactly as
// the reference does (codegen.cpp:68-80), and -- new -- <output_dir>/tinympc/tiny_b200_family.h: the same solver as a plain-C
// initialiser of the batched library's family struct (include/tinympc_b200.h), i.e. the constant tables the sm_100a kernels
// are launched with.  0 on success.
int tiny_codegen(TinySolver* solver, const char* output_dir, int verbose);
// codegen.cpp:82-101: stores the four sensitivity matrices in the cache when adaptive_rho is on, then tiny_codegen
int tiny_codegen_with_sensitivity(TinySolver* solver, const char* output_dir, tinyMatrix* dK, tinyMatrix* dP, tinyMatrix* dC1,
                                  tinyMatrix* dC2, int verbose);
int codegen_create_directories(const char* output_dir, int verbose);
int codegen_data_header(const char* output_dir, int verbose);
int codegen_data_source(TinySolver* solver, const char* output_dir, int verbose);
int codegen_example(const char* output_dir, int verbose);
// new: the family table for the batched C ABI
int codegen_b200_family(TinySolver* solver, const char* output_dir, int verbose);

}  // extern "C"
