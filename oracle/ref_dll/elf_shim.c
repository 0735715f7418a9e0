/*
 * elf_shim.c -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * oracle/_ref/model_simple.so: an ELF face of the reference's dynamics DLL with exactly the symbol set that
 * /root/reference/core/model.py binds on Linux (core/model.py:104-164: `cdll.LoadLibrary` of a per-instance COPY of
 * `core/model_simple.so`, three `void f(void)` functions, ~45 `double` globals through `in_dll`).  With it the
 * reference's OWN Python layers -- core/model.py, core/controller.py, env/ctrl_env.py, unmodified, imported from where
 * they lie -- run in this container on top of the DLL's own machine code (oracle/refpy.py), which is what pins the
 * oracle's env layer (tests/golden/env_golden_refpy.npz).
 *
 * How: the PE image lives in this object's own .bss (`b747_pe_image`, shim_syms.S, generated from the DLL's export
 * directory by gen_shim.py); a constructor lays the embedded DLL out there (pe_host.c: relocations, trapped imports,
 * W^X), and every data export of the DLL is an ELF object symbol at `b747_pe_image + RVA`, so `real_T.in_dll(lib,
 * "state")` resolves straight into the DLL's own global.  Every dlopen of a different COPY of this file gets its own
 * .bss and therefore private globals -- the isolation the reference relies on.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "shim_rva.h"

typedef struct b747ref_inst {
  uint8_t *img; uint32_t img_size, exp_rva, n_rw; uint32_t rw_rva[8], rw_size[8];
} b747ref_inst;
int b747ref_load_at(uint8_t *img, b747ref_inst *in);
uint32_t b747ref_image_size(void);
extern uint8_t b747_pe_image[];
typedef void(__attribute__((ms_abi)) * ms_void_fn)(void);

static b747ref_inst g_inst;

__attribute__((constructor)) static void shim_load(void) {
  if (b747ref_image_size() != B747_SHIM_IMAGE_SIZE || b747ref_load_at(b747_pe_image, &g_inst)) {
    fprintf(stderr, "model_simple.so (oracle shim): cannot lay out the embedded DLL image\n");
    abort();
  }
}

void model_simple_initialize(void) { ((ms_void_fn)(b747_pe_image + B747_SHIM_RVA_model_simple_initialize))(); }
void model_simple_step(void) { ((ms_void_fn)(b747_pe_image + B747_SHIM_RVA_model_simple_step))(); }
void model_simple_terminate(void) { ((ms_void_fn)(b747_pe_image + B747_SHIM_RVA_model_simple_terminate))(); }
