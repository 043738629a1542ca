/*
 * ref_loader.c -- TEST INFRASTRUCTURE (never linked into or imported by the product).
 *
 * Runs the reference's OWN compiled Fortran on this Linux box.  No Fortran compiler exists in the image, but the
 * reference tree ships the object files its last f2py build left behind:
 *     /root/reference/build/temp.win-amd64-3.7/Release/wmf/cuencas.o     (module cu,     cuencas.f90)
 *     /root/reference/build/temp.win-amd64-3.7/Release/wmf/modelosv2.o   (module models, modelosv2.f90)
 * They are x86-64 COFF objects from "GNU Fortran 5.3.0 -O3 -funroll-loops" (no fast-math) whose DWARF line tables end
 * at lines 2701 / 1925 of the 2704 / 1927-line sources in the tree, i.e. they were built from the current sources.
 * oracle/Makefile embeds the two files byte-for-byte (.incbin, straight from /root/reference, nothing is copied into the
 * repository) into oracle/_ref/libwmf_ref.so together with this loader, which
 *   1. maps their .text/.data/.bss/.rdata sections into one RWX region and applies the AMD64 COFF relocations,
 *   2. resolves the few imports (malloc/free/memset/memcpy/realloc, libm's powf/expf/logf/sinf/cosf/tanf/atanf/...,
 *      libgfortran entry points) to ms_abi thunks defined here -- the objects use the Windows x64 calling convention,
 *   3. replaces the routines that do file I/O through libgfortran 5.3's runtime (read_int_basin,
 *      rain_read_ascii_table[_separate], write_float_basin, write_int_basin) by C restatements of those few lines, patched in at
 *      the routines' entry points; `print` statements become no-ops,
 *   4. calls any routine through a generic ms_abi trampoline (Fortran passes everything by reference, so every
 *      argument is a pointer or a hidden character length: integer class).
 * Everything numerical -- every loop of cuencas.f90 / modelosv2.f90 -- is the reference's own machine code; only libm
 * is glibc's instead of msvcrt's.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#define MS __attribute__((ms_abi))

extern const unsigned char wref_blob_cu[], wref_blob_cu_end[];
extern const unsigned char wref_blob_models[], wref_blob_models_end[];

#pragma pack(push, 1)
typedef struct { uint16_t machine, nsect; uint32_t stamp, symptr, nsym; uint16_t optsz, flags; } coff_hdr;
typedef struct { char name[8]; uint32_t vsize, vaddr, rawsz, rawptr, relptr, lineptr; uint16_t nrel, nline; uint32_t flags; } coff_sect;
typedef struct { uint32_t vaddr, sym; uint16_t type; } coff_rel;
typedef struct { union { char s[8]; struct { uint32_t zero, off; } l; } n; uint32_t value; int16_t sect; uint16_t type; uint8_t scl, naux; } coff_sym;
#pragma pack(pop)

typedef struct {
    const unsigned char *blob;
    size_t blob_size;
    const coff_hdr *h;
    const coff_sect *sect;
    const coff_sym *sym;
    const char *strtab;
    unsigned char *region;
    size_t region_size;
    unsigned char **sect_base;  /* per section, NULL when not loaded */
    unsigned char **sym_addr;   /* per symbol index (commons and imports included) */
} wref_obj;

static wref_obj g_obj[2];
static char g_err[512];

const char *wref_error(void) { return g_err; }

/* ------------------------------------------------------------------ imports: ms_abi -> SysV thunks */
/* The reference writes one element past the end of several (k,N_cel) arrays: at the outlet drenaid = N_cel+1 and
 * StoOut(4,drenaid), Fluxes(:,drenaid), Vs/Vsc(:,drena_id) are updated without a guard (modelosv2.f90:468, :617-618,
 * :658-680, :1374, :1601).  On Windows that silently lands in allocator slack; here every block the
 * Fortran allocates gets WREF_PAD spare bytes so glibc's heap checks stay quiet and the stray values go nowhere. */
#define WREF_PAD 256
static MS void *t_malloc(size_t n) { return calloc(1, n + WREF_PAD); }
static MS void t_free(void *p) { free(p); }
static MS void *t_realloc(void *p, size_t n) { return realloc(p, n + WREF_PAD); }
static MS void *t_memcpy(void *d, const void *s, size_t n) { return memcpy(d, s, n); }
static MS void *t_memset(void *d, int c, size_t n) { return memset(d, c, n); }
static MS float t_powf(float a, float b) { return powf(a, b); }
static MS float t_expf(float a) { return expf(a); }
static MS float t_logf(float a) { return logf(a); }
static MS float t_sinf(float a) { return sinf(a); }
static MS float t_cosf(float a) { return cosf(a); }
static MS float t_tanf(float a) { return tanf(a); }
static MS float t_atanf(float a) { return atanf(a); }
static MS float t_coshf(float a) { return coshf(a); }
static MS float t_sinhf(float a) { return sinhf(a); }
static MS float t_tanhf(float a) { return tanhf(a); }
static MS void t_noop(void) {}
static MS void t_os_error(const char *msg) {
    fprintf(stderr, "wref: reference code raised os_error: %s\n", msg ? msg : "?");
    abort();
}
static MS void t_runtime_error_at(const char *where, const char *fmt, ...) {
    fprintf(stderr, "wref: reference code raised a runtime error: %s %s\n", where ? where : "", fmt ? fmt : "");
    abort();
}
static MS void t_unsupported(void) {
    fprintf(stderr, "wref: the reference called a libgfortran routine this loader does not provide\n");
    abort();
}

static const struct { const char *name; void *fn; } k_imports[] = {
    {"malloc", (void *)t_malloc}, {"free", (void *)t_free}, {"realloc", (void *)t_realloc},
    {"memcpy", (void *)t_memcpy}, {"memset", (void *)t_memset},
    {"powf", (void *)t_powf}, {"expf", (void *)t_expf}, {"logf", (void *)t_logf}, {"sinf", (void *)t_sinf},
    {"cosf", (void *)t_cosf}, {"tanf", (void *)t_tanf}, {"atanf", (void *)t_atanf}, {"coshf", (void *)t_coshf},
    {"sinhf", (void *)t_sinhf}, {"tanhf", (void *)t_tanhf},
    {"_gfortran_os_error", (void *)t_os_error}, {"_gfortran_runtime_error_at", (void *)t_runtime_error_at},
    /* `print *` (list-directed writes to stdout): dropped.  The routines that read or write files are replaced. */
    {"_gfortran_st_write", (void *)t_noop}, {"_gfortran_st_write_done", (void *)t_noop},
    {"_gfortran_transfer_character_write", (void *)t_noop}, {"_gfortran_transfer_integer_write", (void *)t_noop},
    {"_gfortran_transfer_real_write", (void *)t_noop}, {"_gfortran_transfer_array_write", (void *)t_noop},
    {"_gfortran_st_open", (void *)t_unsupported}, {"_gfortran_st_close", (void *)t_unsupported},
    {"_gfortran_st_read", (void *)t_unsupported}, {"_gfortran_st_read_done", (void *)t_unsupported},
    {"_gfortran_transfer_array", (void *)t_unsupported}, {"_gfortran_transfer_character", (void *)t_unsupported},
    {"_gfortran_transfer_integer", (void *)t_unsupported}, {"_gfortran_transfer_real", (void *)t_unsupported},
    {"_gfortran_adjustl", (void *)t_unsupported}, {"_gfortran_concat_string", (void *)t_unsupported},
    {"_gfortran_cshift0_4", (void *)t_unsupported},
};

/* ------------------------------------------------------------------ COFF */
static const char *sym_name(const wref_obj *o, const coff_sym *s, char buf[9]) {
    if (s->n.l.zero == 0) return o->strtab + s->n.l.off;
    memcpy(buf, s->n.s, 8);
    buf[8] = 0;
    return buf;
}

static const char *sect_name(const wref_obj *o, const coff_sect *s, char buf[9]) {
    if (s->name[0] == '/') return o->strtab + atoi(s->name + 1);
    memcpy(buf, s->name, 8);
    buf[8] = 0;
    return buf;
}

static int sect_loadable(const wref_obj *o, const coff_sect *s) {
    char b[9];
    const char *n = sect_name(o, s, b);
    if (s->flags & (0x00000800u | 0x00000200u | 0x02000000u)) return 0;          /* LNK_REMOVE, LNK_INFO, MEM_DISCARDABLE */
    if (!strncmp(n, ".debug", 6) || !strcmp(n, ".pdata") || !strcmp(n, ".xdata")) return 0;  /* unwind tables: not needed */
    return s->rawsz > 0;
}

static size_t sect_align(const coff_sect *s) {
    const unsigned a = (s->flags >> 20) & 15;
    return a ? (size_t)1 << (a - 1) : 16;
}

static int load_obj(wref_obj *o, const unsigned char *blob, size_t size) {
    memset(o, 0, sizeof *o);
    o->blob = blob;
    o->blob_size = size;
    o->h = (const coff_hdr *)blob;
    if (size < sizeof(coff_hdr) || o->h->machine != 0x8664) { snprintf(g_err, sizeof g_err, "not an x86-64 COFF object"); return 1; }
    o->sect = (const coff_sect *)(blob + sizeof(coff_hdr) + o->h->optsz);
    o->sym = (const coff_sym *)(blob + o->h->symptr);
    o->strtab = (const char *)(o->sym + o->h->nsym);
    const int ns = o->h->nsect;
    const uint32_t nsym = o->h->nsym;
    /* layout: loadable sections, commons, import thunks */
    size_t total = 0;
    size_t *off = calloc((size_t)ns, sizeof(size_t));
    for (int i = 0; i < ns; ++i) {
        if (!sect_loadable(o, &o->sect[i])) { off[i] = (size_t)-1; continue; }
        const size_t a = sect_align(&o->sect[i]) < 64 ? 64 : sect_align(&o->sect[i]);
        total = (total + a - 1) & ~(a - 1);
        off[i] = total;
        total += o->sect[i].rawsz;
    }
    size_t *soff = calloc(nsym, sizeof(size_t));
    for (uint32_t k = 0; k < nsym; k += 1 + o->sym[k].naux) {
        const coff_sym *s = &o->sym[k];
        soff[k] = (size_t)-1;
        if (s->sect == 0 && s->scl == 2) {
            total = (total + 63) & ~(size_t)63;
            soff[k] = total;
            total += s->value ? s->value : 16;   /* common block of `value` bytes, or a 16-byte import thunk */
        }
    }
    total = (total + 4095) & ~(size_t)4095;
    o->region = mmap(NULL, total, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (o->region == MAP_FAILED) { snprintf(g_err, sizeof g_err, "mmap of %zu RWX bytes failed", total); return 1; }
    o->region_size = total;
    o->sect_base = calloc((size_t)ns, sizeof(void *));
    o->sym_addr = calloc(nsym, sizeof(void *));
    for (int i = 0; i < ns; ++i) {
        if (off[i] == (size_t)-1) continue;
        o->sect_base[i] = o->region + off[i];
        if (o->sect[i].rawptr) memcpy(o->sect_base[i], blob + o->sect[i].rawptr, o->sect[i].rawsz);   /* .bss stays zero */
    }
    for (uint32_t k = 0; k < nsym; k += 1 + o->sym[k].naux) {
        const coff_sym *s = &o->sym[k];
        char b[9];
        if (s->sect > 0) {
            if (o->sect_base[s->sect - 1]) o->sym_addr[k] = o->sect_base[s->sect - 1] + s->value;
        } else if (s->sect == 0 && s->scl == 2) {
            unsigned char *p = o->region + soff[k];
            o->sym_addr[k] = p;
            if (s->value == 0) {  /* import: jmp *0(%rip) ; .quad target */
                const char *n = sym_name(o, s, b);
                void *fn = NULL;
                for (size_t j = 0; j < sizeof k_imports / sizeof k_imports[0]; ++j)
                    if (!strcmp(n, k_imports[j].name)) fn = k_imports[j].fn;
                if (!fn) { snprintf(g_err, sizeof g_err, "unresolved import %s", n); return 1; }
                p[0] = 0xFF; p[1] = 0x25; p[2] = p[3] = p[4] = p[5] = 0;
                memcpy(p + 6, &fn, 8);
            }
        }
    }
    /* relocations */
    for (int i = 0; i < ns; ++i) {
        if (!o->sect_base[i]) continue;
        const coff_sect *sc = &o->sect[i];
        const coff_rel *r = (const coff_rel *)(blob + sc->relptr);
        uint32_t n = sc->nrel;
        if ((sc->flags & 0x01000000u) && n == 0xffff) { n = r[0].vaddr - 1; ++r; }   /* LNK_NRELOC_OVFL */
        for (uint32_t k = 0; k < n; ++k) {
            unsigned char *P = o->sect_base[i] + r[k].vaddr;
            unsigned char *S = o->sym_addr[r[k].sym];
            if (!S) { snprintf(g_err, sizeof g_err, "relocation %u of section %d against an unloaded symbol", k, i); return 1; }
            const unsigned t = r[k].type;
            if (t == 1) {                       /* ADDR64 */
                uint64_t v; memcpy(&v, P, 8); v += (uint64_t)(uintptr_t)S; memcpy(P, &v, 8);
            } else if (t >= 4 && t <= 9) {      /* REL32, REL32_1 .. REL32_5 */
                int32_t a; memcpy(&a, P, 4);
                const int64_t v = (int64_t)a + ((intptr_t)S - ((intptr_t)P + 4 + (int)(t - 4)));
                if (v > INT32_MAX || v < INT32_MIN) { snprintf(g_err, sizeof g_err, "REL32 overflow"); return 1; }
                a = (int32_t)v; memcpy(P, &a, 4);
            } else {
                snprintf(g_err, sizeof g_err, "unsupported relocation type %u in section %d", t, i);
                return 1;
            }
        }
    }
    free(off);
    free(soff);
    return 0;
}

void *wref_sym(int obj, const char *name) {
    if (obj < 0 || obj > 1 || !g_obj[obj].region) return NULL;
    const wref_obj *o = &g_obj[obj];
    for (uint32_t k = 0; k < o->h->nsym; k += 1 + o->sym[k].naux) {
        char b[9];
        if ((o->sym[k].sect > 0 || o->sym[k].value > 0) && !strcmp(sym_name(o, &o->sym[k], b), name)) return o->sym_addr[k];
    }
    return NULL;
}

/* Redirect a routine: mov rax, imm64 ; jmp rax written over its first 12 bytes. */
static int patch(int obj, const char *name, void *target) {
    unsigned char *p = wref_sym(obj, name);
    if (!p) { snprintf(g_err, sizeof g_err, "cannot patch %s: symbol not found", name); return 1; }
    p[0] = 0x48; p[1] = 0xB8;
    memcpy(p + 2, &target, 8);
    p[10] = 0xFF; p[11] = 0xE0;
    return 0;
}

/* ------------------------------------------------------------------ gfortran 5.3 array descriptor (module allocatables) */
typedef struct { int64_t stride, lbound, ubound; } gfc5_dim;
typedef struct { void *base; int64_t offset; int64_t dtype; gfc5_dim dim[2]; } gfc5_desc;

/* Give an allocatable module array of rank 1 or 2 a malloc'ed body holding a copy of `src` (Fortran order), as
 * `allocate(X(n0[,n1])); X = src` would.  type: 1 integer(4), 3 real(4). */
int wref_set_allocatable(void *desc_addr, int rank, int type, int64_t n0, int64_t n1, const void *src) {
    gfc5_desc *d = desc_addr;
    if (!d || rank < 1 || rank > 2) return 1;
    if (d->base) free(d->base);
    const size_t count = (size_t)n0 * (size_t)(rank == 2 ? n1 : 1);
    d->base = malloc(count * 4 + WREF_PAD);
    if (src) memcpy(d->base, src, count * 4); else memset(d->base, 0, count * 4);
    d->dtype = rank | (type << 3) | (4 << 6);
    d->dim[0].stride = 1; d->dim[0].lbound = 1; d->dim[0].ubound = n0;
    d->offset = -1;
    if (rank == 2) {
        d->dim[1].stride = n0; d->dim[1].lbound = 1; d->dim[1].ubound = n1;
        d->offset = -(1 + n0);
    }
    return 0;
}

void wref_free_allocatable(void *desc_addr) {
    gfc5_desc *d = desc_addr;
    if (d && d->base) { free(d->base); d->base = NULL; }
}

/* Read back: base pointer and extents of an allocatable (0 extents when unallocated). */
void *wref_get_allocatable(void *desc_addr, int rank, int64_t *n0, int64_t *n1) {
    gfc5_desc *d = desc_addr;
    *n0 = *n1 = 0;
    if (!d || !d->base) return NULL;
    *n0 = d->dim[0].ubound - d->dim[0].lbound + 1;
    *n1 = rank == 2 ? d->dim[1].ubound - d->dim[1].lbound + 1 : 1;
    return d->base;
}

/* ------------------------------------------------------------------ the replaced file routines (module models) */
static void trim(char *dst, size_t cap, const char *src, int len) {
    int n = len;
    while (n > 0 && (src[n - 1] == ' ' || src[n - 1] == 0)) --n;
    if ((size_t)n >= cap) n = (int)cap - 1;
    memcpy(dst, src, (size_t)n);
    dst[n] = 0;
}

/* read_int_basin, modelosv2.f90:911-926: record `record` of N_cel int32, direct access */
static MS void r_read_int_basin(const char *ruta, const int32_t *record, const int32_t *n_cel, int32_t *vect, int32_t *res, int ruta_len) {
    char path[512];
    trim(path, sizeof path, ruta, ruta_len > 0 && ruta_len <= 500 ? ruta_len : 500);
    *res = 1;
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "wref: read_int_basin cannot open '%s'\n", path); abort(); }
    if (*record >= 1 && fseeko(f, (off_t)4 * *n_cel * (*record - 1), SEEK_SET) == 0 &&
        fread(vect, 4, (size_t)*n_cel, f) == (size_t)*n_cel)
        *res = 0;
    fclose(f);
}

/* write_float_basin / write_int_basin, modelosv2.f90:928-958 */
static void write_record(const char *ruta, int ruta_len, const void *vect, int32_t record, int32_t n_cel, int32_t n_col, int always) {
    if (record <= 0 && !always) return;
    char path[512];
    trim(path, sizeof path, ruta, ruta_len > 0 && ruta_len <= 255 ? ruta_len : 255);
    FILE *f = fopen(path, record == 1 ? "wb" : "r+b");
    if (!f) { fprintf(stderr, "wref: write_*_basin cannot open '%s'\n", path); abort(); }
    fseeko(f, (off_t)4 * n_cel * n_col * (record - 1), SEEK_SET);
    fwrite(vect, 4, (size_t)n_cel * n_col, f);
    fclose(f);
}
static MS void r_write_float_basin(const char *ruta, const float *vect, const int32_t *record, const int32_t *n_cel, const int32_t *n_col, int ruta_len) {
    write_record(ruta, ruta_len, vect, *record, *n_cel, *n_col, 0);
}
static MS void r_write_int_basin(const char *ruta, const int32_t *vect, const int32_t *record, const int32_t *n_cel, const int32_t *n_col, int ruta_len) {
    write_record(ruta, ruta_len, vect, *record, *n_cel, *n_col, 1);
}

/* rain_read_ascii_table, modelosv2.f90:966-1002: allocate idEvento / posEvento(Nintervals) and fill them from the .hdr
 * (three "a b c Ntotal" lines, three skipped lines, rain_first_point skipped lines when > 1, then "id, pos, ..." rows) */
static void next_line(FILE *f, char *buf, size_t cap) { if (!fgets(buf, (int)cap, f)) buf[0] = 0; }
static MS void r_rain_read_ascii_table(const char *ruta, const int32_t *nintervals, int ruta_len) {
    char path[512], line[1024];
    trim(path, sizeof path, ruta, ruta_len > 0 && ruta_len <= 500 ? ruta_len : 255);
    const int n = *nintervals;
    int32_t *ids = calloc((size_t)(n > 0 ? n : 1), 4), *pos = calloc((size_t)(n > 0 ? n : 1), 4);
    FILE *f = fopen(path, "r");
    if (!f) { fprintf(stderr, "wref: rain_read_ascii_table cannot open '%s'\n", path); abort(); }
    long ntotal = 0;
    for (int i = 0; i < 3; ++i) {  /* read(10,*) oe, oe, oe, Ntotal */
        next_line(f, line, sizeof line);
        char *tok[4] = {0};
        int k = 0;
        for (char *t = strtok(line, " ,\t\r\n"); t && k < 4; t = strtok(NULL, " ,\t\r\n")) tok[k++] = t;
        if (k == 4) ntotal = strtol(tok[3], NULL, 10);
    }
    for (int i = 0; i < 3; ++i) next_line(f, line, sizeof line);
    const int32_t rfp = *(int32_t *)wref_sym(1, "__models_MOD_rain_first_point");
    if (n + rfp <= ntotal + 1) {
        if (rfp > 1) for (int i = 0; i < rfp; ++i) next_line(f, line, sizeof line);
        for (int i = 0; i < n; ++i) {
            next_line(f, line, sizeof line);
            char *t = strtok(line, " ,\t\r\n");
            if (t) ids[i] = (int32_t)strtol(t, NULL, 10);
            t = strtok(NULL, " ,\t\r\n");
            if (t) pos[i] = (int32_t)strtol(t, NULL, 10);
        }
    }
    fclose(f);
    wref_set_allocatable(wref_sym(1, "__models_MOD_idevento"), 1, 1, n, 1, ids);
    wref_set_allocatable(wref_sym(1, "__models_MOD_posevento"), 1, 1, n, 1, pos);
    free(ids);
    free(pos);
}

/* rain_read_ascii_table_separate, modelosv2.f90:1004-1065: posConv / posStra(Nintervals) from the two headers; note the
 * stricter test Nintervals + rain_first_point <= Ntotal (the plain table reader allows Ntotal + 1) */
static void read_pos_column(const char *ruta, int ruta_len, int n, int32_t *pos) {
    char path[512], line[1024];
    trim(path, sizeof path, ruta, ruta_len > 0 && ruta_len <= 500 ? ruta_len : 255);
    FILE *f = fopen(path, "r");
    if (!f) { fprintf(stderr, "wref: rain_read_ascii_table_separate cannot open '%s'\n", path); abort(); }
    long ntotal = 0;
    for (int i = 0; i < 3; ++i) {
        next_line(f, line, sizeof line);
        char *tok[4] = {0};
        int k = 0;
        for (char *t = strtok(line, " ,\t\r\n"); t && k < 4; t = strtok(NULL, " ,\t\r\n")) tok[k++] = t;
        if (k == 4) ntotal = strtol(tok[3], NULL, 10);
    }
    for (int i = 0; i < 3; ++i) next_line(f, line, sizeof line);
    const int32_t rfp = *(int32_t *)wref_sym(1, "__models_MOD_rain_first_point");
    if (n + rfp <= ntotal) {
        if (rfp > 1) for (int i = 0; i < rfp; ++i) next_line(f, line, sizeof line);
        for (int i = 0; i < n; ++i) {
            next_line(f, line, sizeof line);
            char *t = strtok(line, " ,\t\r\n");
            t = strtok(NULL, " ,\t\r\n");
            if (t) pos[i] = (int32_t)strtol(t, NULL, 10);
        }
    }
    fclose(f);
}
static MS void r_rain_read_ascii_table_separate(const char *ruta_conv, const char *ruta_stra, const int32_t *nintervals,
                                                int len_conv, int len_stra) {
    const int n = *nintervals;
    int32_t *pc = calloc((size_t)(n > 0 ? n : 1), 4), *ps = calloc((size_t)(n > 0 ? n : 1), 4);
    read_pos_column(ruta_conv, len_conv, n, pc);
    read_pos_column(ruta_stra, len_stra, n, ps);
    wref_set_allocatable(wref_sym(1, "__models_MOD_posconv"), 1, 1, n, 1, pc);
    wref_set_allocatable(wref_sym(1, "__models_MOD_posstra"), 1, 1, n, 1, ps);
    free(pc);
    free(ps);
}

/* ------------------------------------------------------------------ entry points (SysV, called through ctypes) */
int wref_init(void) {
    if (g_obj[0].region) return 0;
    if (load_obj(&g_obj[0], wref_blob_cu, (size_t)(wref_blob_cu_end - wref_blob_cu))) return 1;
    if (load_obj(&g_obj[1], wref_blob_models, (size_t)(wref_blob_models_end - wref_blob_models))) return 1;
    if (patch(1, "__models_MOD_read_int_basin", (void *)r_read_int_basin)) return 1;
    if (patch(1, "__models_MOD_write_float_basin", (void *)r_write_float_basin)) return 1;
    if (patch(1, "__models_MOD_write_int_basin", (void *)r_write_int_basin)) return 1;
    if (patch(1, "__models_MOD_rain_read_ascii_table", (void *)r_rain_read_ascii_table)) return 1;
    if (patch(1, "__models_MOD_rain_read_ascii_table_separate", (void *)r_rain_read_ascii_table_separate)) return 1;
    return 0;
}

typedef uint64_t u64;
typedef void (MS *msfn)(u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64,
                        u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64,
                        u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64, u64);

/* Call a Fortran subroutine with up to 48 integer-class arguments (pointers and hidden character lengths). */
int wref_call(void *fn, int nargs, const u64 *args) {
    if (!fn || nargs < 0 || nargs > 48) return 1;
    u64 a[48] = {0};
    memcpy(a, args, sizeof(u64) * (size_t)nargs);
    ((msfn)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15],
               a[16], a[17], a[18], a[19], a[20], a[21], a[22], a[23], a[24], a[25], a[26], a[27], a[28], a[29], a[30], a[31],
               a[32], a[33], a[34], a[35], a[36], a[37], a[38], a[39], a[40], a[41], a[42], a[43], a[44], a[45], a[46], a[47]);
    return 0;
}
