/*
 * pe_host.c -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Runs the reference's own dynamics library -- the Simulink-Coder DLL that
 * core/model.py:104-126 loads (core/model_simple_win64.dll) -- natively on an
 * x86-64 Linux host.  The DLL image is embedded into this shared object at
 * build time (oracle/Makefile, `.incbin` of the file where it lies under
 * /root/reference); nothing of the reference is copied into the repository.
 *
 * How: a PE32+ image is just code + data.  Each `b747ref_open()` maps a private
 * copy of the sections, applies the base relocations (so any number of
 * independent instances can live in one process -- the reference gets the
 * same isolation by copying the DLL file per Model, core/model.py:99-110),
 * points every import-table slot at a trap (the numeric path never calls
 * into KERNEL32), and resolves the export directory.  DllMain / CRT start-up
 * is deliberately NOT run, so the statically linked UCRT libm keeps its
 * file-default ISA flags (SSE2 path).  Exports are `void f(void)` and are
 * called through the Microsoft x64 ABI.
 *
 * The DLL is an unaudited binary from the reference tree, so: the build refuses any file whose SHA-256 differs from the
 * pinned one (oracle/Makefile, ref_dll/model_simple_win64.sha256); pages are never writable and executable at once
 * (sections are laid out read-write, then code is re-protected read + execute and data loses execute); and the worker
 * processes that run it in bulk drop network, exec, ptrace and file-write syscalls first (b747ref_sandbox, seccomp-bpf).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use this file's output (oracle/_ref/).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <stddef.h>
#include <errno.h>
#include <fcntl.h>
#include <linux/audit.h>
#include <linux/filter.h>
#include <linux/seccomp.h>
#include <sys/prctl.h>
#include <sys/syscall.h>

#ifndef B747_DLL_PATH
#error "build with -DB747_DLL_PATH=\"/root/reference/core/model_simple_win64.dll\""
#endif

__asm__(
    ".section .rodata\n"
    ".balign 16\n"
    ".globl b747ref_image_begin\n"
    "b747ref_image_begin:\n"
    ".incbin \"" B747_DLL_PATH "\"\n"
    ".globl b747ref_image_end\n"
    "b747ref_image_end:\n"
    ".previous\n");
extern const uint8_t b747ref_image_begin[], b747ref_image_end[];

typedef void(__attribute__((ms_abi)) * ms_void_fn)(void);

typedef struct b747ref_inst {
  uint8_t *img;      /* mapped image */
  uint32_t img_size; /* SizeOfImage */
  uint32_t exp_rva;  /* export directory RVA */
  /* writable sections (for snapshot/restore) */
  uint32_t n_rw;
  uint32_t rw_rva[8], rw_size[8];
} b747ref_inst;

static void __attribute__((ms_abi)) import_trap(void) {
  fprintf(stderr, "b747ref: the DLL called into an import (Win32/CRT path) -- not supported\n");
  abort();
}

/* The legacy model_win64.dll (MinGW) links its CRT dynamically: the numeric path reaches msvcrt for `asin` and the
 * mem* / allocation routines (everything else -- sin, cos, atan2, pow, exp, sqrt -- is MinGW's static libmingwex inside the
 * DLL).  These are resolved by NAME to host functions behind Microsoft-ABI thunks; every other import stays a trap. */
#include <math.h>
static double __attribute__((ms_abi)) t_asin(double x) { return asin(x); }
static void *__attribute__((ms_abi)) t_memcpy(void *d, const void *s_, size_t n) { return memcpy(d, s_, n); }
static void *__attribute__((ms_abi)) t_memset(void *d, int c, size_t n) { return memset(d, c, n); }
static void *__attribute__((ms_abi)) t_malloc(size_t n) { return malloc(n); }
static void *__attribute__((ms_abi)) t_calloc(size_t a, size_t b) { return calloc(a, b); }
static void __attribute__((ms_abi)) t_free(void *p) { free(p); }
static const struct { const char *name; void *fn; } k_thunks[] = {
    {"asin", (void *)t_asin}, {"memcpy", (void *)t_memcpy}, {"memset", (void *)t_memset},
    {"malloc", (void *)t_malloc}, {"calloc", (void *)t_calloc}, {"free", (void *)t_free},
};
static void *resolve_import(const char *name) {
  for (size_t i = 0; i < sizeof k_thunks / sizeof k_thunks[0]; i++)
    if (!strcmp(k_thunks[i].name, name)) return k_thunks[i].fn;
  return (void *)import_trap;
}

static inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

/* Lay the embedded PE image out at `img` (page aligned, at least b747ref_image_size() bytes, readable + writable,
 * zero-filled), apply the base relocations for that address, trap the imports, then tighten the page protections per
 * section: code becomes read + execute, read-only data read-only, nothing is ever writable and executable at once. */
static int load_image(uint8_t *img, b747ref_inst *in) {
  const uint8_t *file = b747ref_image_begin;
  size_t file_size = (size_t)(b747ref_image_end - b747ref_image_begin);
  if (file_size < 0x400 || rd16(file) != 0x5a4d) return -1;
  uint32_t nt = rd32(file + 0x3c);
  if (rd32(file + nt) != 0x00004550) return -1;
  const uint8_t *fh = file + nt + 4;      /* IMAGE_FILE_HEADER */
  uint16_t n_sections = rd16(fh + 2);
  uint16_t opt_size = rd16(fh + 16);
  const uint8_t *opt = fh + 20;           /* IMAGE_OPTIONAL_HEADER64 */
  if (rd16(opt) != 0x20b) return -1;      /* PE32+ only */
  uint64_t preferred = rd64(opt + 24);
  uint32_t img_size = rd32(opt + 56);
  uint32_t hdr_size = rd32(opt + 60);
  const uint8_t *dirs = opt + 112;        /* data directories */
  uint32_t exp_rva = rd32(dirs + 0 * 8);
  uint32_t imp_rva = rd32(dirs + 1 * 8);
  uint32_t rel_rva = rd32(dirs + 5 * 8), rel_size = rd32(dirs + 5 * 8 + 4);
  in->img = img; in->img_size = img_size; in->exp_rva = exp_rva;

  memcpy(img, file, hdr_size);
  const uint8_t *sec = opt + opt_size;
  for (unsigned i = 0; i < n_sections; i++, sec += 40) {
    uint32_t vsize = rd32(sec + 8), va = rd32(sec + 12), raw_size = rd32(sec + 16), raw_off = rd32(sec + 20);
    uint32_t flags = rd32(sec + 36);
    if (raw_size) memcpy(img + va, file + raw_off, raw_size < vsize || !vsize ? raw_size : vsize);
    if ((flags & 0x80000000u) && in->n_rw < 8) { /* IMAGE_SCN_MEM_WRITE */
      in->rw_rva[in->n_rw] = va;
      in->rw_size[in->n_rw] = vsize ? vsize : raw_size;
      in->n_rw++;
    }
  }
  /* base relocations: blocks of {page_rva, block_size, uint16 entries[]}; type 10 = DIR64 */
  int64_t delta = (int64_t)((uint64_t)img - preferred);
  if (delta && rel_size) {
    uint8_t *p = img + rel_rva, *end = p + rel_size;
    while (p + 8 <= end) {
      uint32_t page = rd32(p), bsz = rd32(p + 4);
      if (bsz < 8) break;
      for (uint32_t k = 8; k + 2 <= bsz; k += 2) {
        uint16_t e = rd16(p + k);
        if ((e >> 12) == 10) {
          uint8_t *slot = img + page + (e & 0xfff);
          uint64_t v = rd64(slot) + (uint64_t)delta;
          memcpy(slot, &v, 8);
        }
      }
      p += bsz;
    }
  }
  /* import table: every IAT slot -> trap, except the few msvcrt routines resolved by name (resolve_import) */
  if (imp_rva) {
    for (uint8_t *d = img + imp_rva; rd32(d + 12); d += 20) {
      uint64_t *iat = (uint64_t *)(img + rd32(d + 16));
      const uint64_t *names = rd32(d) ? (const uint64_t *)(img + rd32(d)) : NULL; /* OriginalFirstThunk: hint/name RVAs */
      for (unsigned k = 0; iat[k]; k++) {
        void *fn = (void *)import_trap;
        if (names && names[k] && !(names[k] >> 63) && names[k] < img_size) fn = resolve_import((const char *)img + names[k] + 2);
        iat[k] = (uint64_t)(uintptr_t)fn;
      }
    }
  }
  /* W^X: headers and read-only sections -> R, code -> R+X, writable data stays RW (never executable) */
  mprotect(img, (hdr_size + 0xfff) & ~0xfffu, PROT_READ);
  sec = opt + opt_size;
  for (unsigned i = 0; i < n_sections; i++, sec += 40) {
    uint32_t vsize = rd32(sec + 8), va = rd32(sec + 12), raw_size = rd32(sec + 16), flags = rd32(sec + 36);
    uint32_t len = ((vsize ? vsize : raw_size) + 0xfff) & ~0xfffu;
    int prot = PROT_READ;
    if (flags & 0x20000000u) prot |= PROT_EXEC;        /* IMAGE_SCN_MEM_EXECUTE */
    else if (flags & 0x80000000u) prot |= PROT_WRITE;  /* IMAGE_SCN_MEM_WRITE (a section that is both stays R+X) */
    if (mprotect(img + va, len, prot)) return -2;
  }
  return 0;
}

uint32_t b747ref_image_size(void) {
  const uint8_t *file = b747ref_image_begin;
  uint32_t nt = rd32(file + 0x3c);
  return rd32(file + nt + 4 + 20 + 56);
}

/* for the ELF face of the DLL (elf_shim.c): the image lives in the shim's own .bss */
int b747ref_load_at(uint8_t *img, b747ref_inst *in) {
  memset(in, 0, sizeof *in);
  return load_image(img, in);
}

b747ref_inst *b747ref_open(void) {
  uint32_t img_size = b747ref_image_size();
  uint8_t *img = mmap(NULL, img_size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (img == MAP_FAILED) return NULL;
  b747ref_inst *in = calloc(1, sizeof *in);
  if (load_image(img, in)) { munmap(img, img_size); free(in); return NULL; }
  return in;
}

void b747ref_close(b747ref_inst *in) {
  if (!in) return;
  munmap(in->img, in->img_size);
  free(in);
}

void *b747ref_sym(b747ref_inst *in, const char *name) {
  const uint8_t *e = in->img + in->exp_rva;
  uint32_t n_names = rd32(e + 24);
  const uint32_t *funcs = (const uint32_t *)(in->img + rd32(e + 28));
  const uint32_t *names = (const uint32_t *)(in->img + rd32(e + 32));
  const uint16_t *ords = (const uint16_t *)(in->img + rd32(e + 36));
  for (uint32_t i = 0; i < n_names; i++)
    if (!strcmp((const char *)in->img + names[i], name)) return in->img + funcs[ords[i]];
  return NULL;
}

void *b747ref_rva(b747ref_inst *in, uint32_t rva) { return in->img + rva; }

void b747ref_call(void *fn) { ((ms_void_fn)fn)(); }

void b747ref_call_n(void *fn, long n) {
  ms_void_fn f = (ms_void_fn)fn;
  for (long i = 0; i < n; i++) f();
}

/* snapshot / restore of all writable sections (lets one instance multiplex envs) */
size_t b747ref_state_size(b747ref_inst *in) {
  size_t s = 0;
  for (uint32_t i = 0; i < in->n_rw; i++) s += in->rw_size[i];
  return s;
}
void b747ref_state_save(b747ref_inst *in, void *buf) {
  uint8_t *b = buf;
  for (uint32_t i = 0; i < in->n_rw; i++) { memcpy(b, in->img + in->rw_rva[i], in->rw_size[i]); b += in->rw_size[i]; }
}
void b747ref_state_load(b747ref_inst *in, const void *buf) {
  const uint8_t *b = buf;
  for (uint32_t i = 0; i < in->n_rw; i++) { memcpy(in->img + in->rw_rva[i], b, in->rw_size[i]); b += in->rw_size[i]; }
}

/* Worker-process sandbox (bench.py's reference-arm / cpu_baseline workers and the golden generators call it before they
 * run the DLL): the import table is trapped, but unaudited machine code could still issue raw syscalls, so the process
 * gives up -- irrevocably, seccomp-bpf -- everything a numeric loop does not need: no new sockets or connections, no
 * exec, no ptrace / process_vm access, no module / bpf / mount calls, and files can only be opened read-only.  Pipes
 * that are already open (the multiprocessing result channel) keep working.  Returns 0 on success. */
#define DENY(nr) BPF_JUMP(BPF_JMP | BPF_JEQ | BPF_K, (nr), 0, 1), BPF_STMT(BPF_RET | BPF_K, SECCOMP_RET_ERRNO | EPERM)
int b747ref_sandbox(void) {
  struct sock_filter f[] = {
      BPF_STMT(BPF_LD | BPF_W | BPF_ABS, offsetof(struct seccomp_data, arch)),
      BPF_JUMP(BPF_JMP | BPF_JEQ | BPF_K, AUDIT_ARCH_X86_64, 1, 0),
      BPF_STMT(BPF_RET | BPF_K, SECCOMP_RET_KILL_PROCESS),
      BPF_STMT(BPF_LD | BPF_W | BPF_ABS, offsetof(struct seccomp_data, nr)),
      DENY(SYS_socket), DENY(SYS_connect), DENY(SYS_bind), DENY(SYS_listen), DENY(SYS_accept), DENY(SYS_accept4),
      DENY(SYS_execve), DENY(SYS_execveat), DENY(SYS_ptrace), DENY(SYS_process_vm_readv), DENY(SYS_process_vm_writev),
      DENY(SYS_mount), DENY(SYS_umount2), DENY(SYS_pivot_root), DENY(SYS_chroot), DENY(SYS_init_module),
      DENY(SYS_finit_module), DENY(SYS_delete_module), DENY(SYS_kexec_load), DENY(SYS_bpf), DENY(SYS_perf_event_open),
      DENY(SYS_unlink), DENY(SYS_unlinkat), DENY(SYS_rename), DENY(SYS_renameat), DENY(SYS_renameat2), DENY(SYS_creat),
      DENY(SYS_truncate), DENY(SYS_chmod), DENY(SYS_fchmodat), DENY(SYS_chown), DENY(SYS_fchownat), DENY(SYS_link),
      DENY(SYS_linkat), DENY(SYS_symlink), DENY(SYS_symlinkat), DENY(SYS_mkdir), DENY(SYS_mkdirat), DENY(SYS_rmdir),
      DENY(SYS_setuid), DENY(SYS_setgid), DENY(SYS_setreuid), DENY(SYS_setregid), DENY(SYS_setresuid), DENY(SYS_setresgid),
      DENY(SYS_kill), DENY(SYS_tkill), DENY(SYS_tgkill),
      /* open(path, flags) / openat(dirfd, path, flags): read-only opens only */
      BPF_JUMP(BPF_JMP | BPF_JEQ | BPF_K, SYS_open, 0, 4),
      BPF_STMT(BPF_LD | BPF_W | BPF_ABS, offsetof(struct seccomp_data, args[1])),
      BPF_JUMP(BPF_JMP | BPF_JSET | BPF_K, O_WRONLY | O_RDWR | O_CREAT | O_TRUNC | O_APPEND, 0, 1),
      BPF_STMT(BPF_RET | BPF_K, SECCOMP_RET_ERRNO | EPERM),
      BPF_STMT(BPF_RET | BPF_K, SECCOMP_RET_ALLOW),
      BPF_JUMP(BPF_JMP | BPF_JEQ | BPF_K, SYS_openat, 0, 4),
      BPF_STMT(BPF_LD | BPF_W | BPF_ABS, offsetof(struct seccomp_data, args[2])),
      BPF_JUMP(BPF_JMP | BPF_JSET | BPF_K, O_WRONLY | O_RDWR | O_CREAT | O_TRUNC | O_APPEND, 0, 1),
      BPF_STMT(BPF_RET | BPF_K, SECCOMP_RET_ERRNO | EPERM),
      BPF_STMT(BPF_RET | BPF_K, SECCOMP_RET_ALLOW),
      BPF_STMT(BPF_RET | BPF_K, SECCOMP_RET_ALLOW),
  };
  struct sock_fprog prog = {(unsigned short)(sizeof f / sizeof f[0]), f};
  if (prctl(PR_SET_NO_NEW_PRIVS, 1, 0, 0, 0)) return -1;
  if (prctl(PR_SET_SECCOMP, SECCOMP_MODE_FILTER, &prog, 0, 0)) return -2;
  return 0;
}
#undef DENY
