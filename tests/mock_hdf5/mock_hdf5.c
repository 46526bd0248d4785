/* A stand-in for libhdf5 (test infrastructure only): implements the ~30 entry points ppkmhd_b200/host/IO_HDF5.cpp binds at
 * run time, with the semantics those calls rely on -- files, datasets written / read through a memory-space hyperslab
 * selection, scalar and string attributes -- on top of a trivial container (a directory "<name>.h5" holding one raw file per
 * dataset / attribute). It lets the tests drive the real writer / reader / restart path (call sequence, dataset and attribute
 * names, dimensions, start / count of the selections) in an image that has no HDF5. It is NOT the HDF5 file format.
 *   gcc -shared -fPIC -O1 -o libmockhdf5.so mock_hdf5.c        PPK_HDF5_LIB=/path/libmockhdf5.so */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

typedef int64_t hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;

hid_t H5T_NATIVE_DOUBLE_g = -1, H5T_NATIVE_INT_g = -1, H5T_C_S1_g = -1, H5P_CLS_DATASET_CREATE_ID_g = -1;

enum { K_FREE = 0, K_FILE, K_SPACE, K_PLIST, K_DSET, K_ATTR, K_TYPE, K_GROUP };
typedef struct {
  int kind, rank, selected, tsize, is_vlen_str, readonly;
  hsize_t dims[3], start[3], count[3], chunk[3];
  char path[512];
} obj_t;
static obj_t objs[256];

static hid_t alloc_obj(int kind) {
  for (int i = 16; i < 256; ++i)
    if (objs[i].kind == K_FREE) {
      memset(&objs[i], 0, sizeof(obj_t));
      objs[i].kind = kind;
      return i;
    }
  return -1;
}
static obj_t *get(hid_t id, int kind) { return (id >= 16 && id < 256 && objs[id].kind == kind) ? &objs[id] : NULL; }

herr_t H5open(void) {
  H5T_NATIVE_DOUBLE_g = 1; H5T_NATIVE_INT_g = 2; H5T_C_S1_g = 3; H5P_CLS_DATASET_CREATE_ID_g = 4;
  return 0;
}
herr_t H5get_libversion(unsigned *a, unsigned *b, unsigned *c) { *a = 1; *b = 14; *c = 0; return 0; }

hid_t H5Fcreate(const char *name, unsigned flags, hid_t fcpl, hid_t fapl) {
  (void)fcpl; (void)fapl;
  if (flags != 2u) return -1; /* H5F_ACC_TRUNC */
  hid_t id = alloc_obj(K_FILE);
  if (id < 0) return -1;
  snprintf(objs[id].path, sizeof(objs[id].path), "%s", name);
  remove(name);
  if (mkdir(name, 0755) != 0) { struct stat sb; if (stat(name, &sb) != 0 || !S_ISDIR(sb.st_mode)) { objs[id].kind = K_FREE; return -1; } }
  return id;
}
hid_t H5Fopen(const char *name, unsigned flags, hid_t fapl) {
  (void)fapl;
  struct stat sb;
  if (flags != 0u || stat(name, &sb) != 0 || !S_ISDIR(sb.st_mode)) return -1; /* H5F_ACC_RDONLY */
  hid_t id = alloc_obj(K_FILE);
  if (id < 0) return -1;
  snprintf(objs[id].path, sizeof(objs[id].path), "%s", name);
  objs[id].readonly = 1;
  return id;
}
herr_t H5Fflush(hid_t f, int scope) { (void)scope; return get(f, K_FILE) ? 0 : -1; }
herr_t H5Fclose(hid_t f) { obj_t *o = get(f, K_FILE); if (!o) return -1; o->kind = K_FREE; return 0; }

hid_t H5Screate_simple(int rank, const hsize_t *dims, const hsize_t *maxdims) {
  (void)maxdims;
  if (rank < 1 || rank > 3) return -1;
  hid_t id = alloc_obj(K_SPACE);
  if (id < 0) return -1;
  objs[id].rank = rank;
  for (int d = 0; d < rank; ++d) { objs[id].dims[d] = dims[d]; objs[id].start[d] = 0; objs[id].count[d] = dims[d]; }
  return id;
}
hid_t H5Screate(int cls) { if (cls != 0) return -1; hid_t id = alloc_obj(K_SPACE); if (id >= 0) objs[id].rank = 0; return id; }
herr_t H5Sselect_hyperslab(hid_t s, int op, const hsize_t *start, const hsize_t *stride, const hsize_t *count, const hsize_t *block) {
  obj_t *o = get(s, K_SPACE);
  if (!o || op != 0) return -1;
  for (int d = 0; d < o->rank; ++d) {
    if ((stride && stride[d] != 1) || (block && block[d] != 1) || start[d] + count[d] > o->dims[d]) return -1;
    o->start[d] = start[d]; o->count[d] = count[d];
  }
  o->selected = 1;
  return 0;
}
herr_t H5Sclose(hid_t s) { obj_t *o = get(s, K_SPACE); if (!o) return -1; o->kind = K_FREE; return 0; }

hid_t H5Pcreate(hid_t cls) { return cls == H5P_CLS_DATASET_CREATE_ID_g ? alloc_obj(K_PLIST) : -1; }
herr_t H5Pset_chunk(hid_t p, int rank, const hsize_t *dims) {
  obj_t *o = get(p, K_PLIST);
  if (!o || rank < 1 || rank > 3) return -1;
  o->rank = rank;
  for (int d = 0; d < rank; ++d) { if (dims[d] == 0) return -1; o->chunk[d] = dims[d]; }
  return 0;
}
herr_t H5Pset_shuffle(hid_t p) { return get(p, K_PLIST) ? 0 : -1; }
herr_t H5Pset_deflate(hid_t p, unsigned level) { return get(p, K_PLIST) && level <= 9 ? 0 : -1; }
herr_t H5Pclose(hid_t p) { obj_t *o = get(p, K_PLIST); if (!o) return -1; o->kind = K_FREE; return 0; }

static int space_elems(const obj_t *s) { long n = 1; for (int d = 0; d < s->rank; ++d) n *= (long)s->count[d]; return (int)n; }

hid_t H5Dcreate2(hid_t f, const char *name, hid_t type, hid_t space, hid_t lcpl, hid_t dcpl, hid_t dapl) {
  (void)lcpl; (void)dapl;
  obj_t *fo = get(f, K_FILE), *so = get(space, K_SPACE), *po = get(dcpl, K_PLIST);
  if (!fo || fo->readonly || !so || type != H5T_NATIVE_DOUBLE_g || name[0] != '/') return -1;
  if (po && po->rank) { /* a chunk may not exceed the dataset (HDF5 rejects that for fixed-size datasets) */
    if (po->rank != so->rank) return -1;
    for (int d = 0; d < so->rank; ++d) if (po->chunk[d] > so->dims[d]) return -1;
  }
  hid_t id = alloc_obj(K_DSET);
  if (id < 0) return -1;
  fo = get(f, K_FILE); so = get(space, K_SPACE);
  snprintf(objs[id].path, sizeof(objs[id].path), "%s/%s.dset", fo->path, name + 1);
  objs[id].rank = so->rank;
  memcpy(objs[id].dims, so->dims, sizeof(so->dims));
  return id;
}
hid_t H5Dopen2(hid_t f, const char *name, hid_t dapl) {
  (void)dapl;
  obj_t *fo = get(f, K_FILE);
  if (!fo || name[0] != '/') return -1;
  char path[512];
  snprintf(path, sizeof(path), "%s/%s.dset", fo->path, name + 1);
  FILE *fp = fopen(path, "rb");
  if (!fp) return -1;
  hid_t id = alloc_obj(K_DSET);
  if (id < 0) { fclose(fp); return -1; }
  int rank = 0;
  if (fread(&rank, sizeof(int), 1, fp) != 1 || rank < 1 || rank > 3 || fread(objs[id].dims, sizeof(hsize_t), 3, fp) != 3) { fclose(fp); objs[id].kind = K_FREE; return -1; }
  fclose(fp);
  objs[id].rank = rank;
  snprintf(objs[id].path, sizeof(objs[id].path), "%s", path);
  return id;
}
/* copy between the selected region of a row-major memory array and a contiguous file array */
static void copy_sel(const obj_t *ms, double *mem, double *file, int to_file) {
  hsize_t c[3] = {1, 1, 1}, s[3] = {0, 0, 0}, d[3] = {1, 1, 1};
  const int off = 3 - ms->rank;
  for (int k = 0; k < ms->rank; ++k) { c[off + k] = ms->count[k]; s[off + k] = ms->start[k]; d[off + k] = ms->dims[k]; }
  size_t n = 0;
  for (hsize_t a = 0; a < c[0]; ++a)
    for (hsize_t b = 0; b < c[1]; ++b)
      for (hsize_t e = 0; e < c[2]; ++e, ++n) {
        const size_t m = ((s[0] + a) * d[1] + (s[1] + b)) * d[2] + (s[2] + e);
        if (to_file) file[n] = mem[m]; else mem[m] = file[n];
      }
}
herr_t H5Dwrite(hid_t ds, hid_t type, hid_t mspace, hid_t fspace, hid_t xfer, const void *buf) {
  (void)xfer;
  obj_t *d = get(ds, K_DSET), *ms = get(mspace, K_SPACE), *fs = get(fspace, K_SPACE);
  if (!d || !ms || !fs || type != H5T_NATIVE_DOUBLE_g || ms->rank != fs->rank || space_elems(ms) != space_elems(fs)) return -1;
  const size_t n = (size_t)space_elems(fs);
  double *tmp = (double *)malloc(n * sizeof(double));
  copy_sel(ms, (double *)buf, tmp, 1);
  FILE *fp = fopen(d->path, "wb");
  if (!fp) { free(tmp); return -1; }
  fwrite(&d->rank, sizeof(int), 1, fp);
  fwrite(d->dims, sizeof(hsize_t), 3, fp);
  fwrite(tmp, sizeof(double), n, fp);
  fclose(fp);
  free(tmp);
  return 0;
}
herr_t H5Dread(hid_t ds, hid_t type, hid_t mspace, hid_t fspace, hid_t xfer, void *buf) {
  (void)xfer;
  obj_t *d = get(ds, K_DSET), *ms = get(mspace, K_SPACE), *fs = get(fspace, K_SPACE);
  if (!d || !ms || !fs || type != H5T_NATIVE_DOUBLE_g || ms->rank != fs->rank || space_elems(ms) != space_elems(fs)) return -1;
  for (int k = 0; k < fs->rank; ++k) if (fs->dims[k] != d->dims[k]) return -1; /* the caller's file space must match the dataset */
  const size_t n = (size_t)space_elems(fs);
  double *tmp = (double *)malloc(n * sizeof(double));
  FILE *fp = fopen(d->path, "rb");
  if (!fp) { free(tmp); return -1; }
  fseek(fp, sizeof(int) + 3 * sizeof(hsize_t), SEEK_SET);
  const size_t got = fread(tmp, sizeof(double), n, fp);
  fclose(fp);
  if (got != n) { free(tmp); return -1; }
  copy_sel(ms, (double *)buf, tmp, 0);
  free(tmp);
  return 0;
}
herr_t H5Dclose(hid_t ds) { obj_t *o = get(ds, K_DSET); if (!o) return -1; o->kind = K_FREE; return 0; }

static const char *owner_path(hid_t loc) {
  obj_t *o = get(loc, K_FILE);
  if (!o) o = get(loc, K_GROUP);
  return o ? o->path : NULL;
}
hid_t H5Acreate2(hid_t loc, const char *name, hid_t type, hid_t space, hid_t acpl, hid_t aapl) {
  (void)acpl; (void)aapl;
  const char *base = owner_path(loc);
  if (!base || !get(space, K_SPACE)) return -1;
  hid_t id = alloc_obj(K_ATTR);
  if (id < 0) return -1;
  snprintf(objs[id].path, sizeof(objs[id].path), "%s/%s.attr", owner_path(loc), name);
  obj_t *t = get(type, K_TYPE);
  objs[id].is_vlen_str = t ? t->is_vlen_str : 0;
  objs[id].tsize = type == H5T_NATIVE_DOUBLE_g ? 8 : (type == H5T_NATIVE_INT_g ? 4 : 0);
  return id;
}
hid_t H5Aopen(hid_t loc, const char *name, hid_t aapl) {
  (void)aapl;
  const char *base = owner_path(loc);
  if (!base) return -1;
  char path[512];
  snprintf(path, sizeof(path), "%s/%s.attr", base, name);
  FILE *fp = fopen(path, "rb");
  if (!fp) return -1;
  fclose(fp);
  hid_t id = alloc_obj(K_ATTR);
  if (id >= 0) snprintf(objs[id].path, sizeof(objs[id].path), "%s", path);
  return id;
}
herr_t H5Awrite(hid_t a, hid_t type, const void *buf) {
  obj_t *o = get(a, K_ATTR);
  if (!o) return -1;
  FILE *fp = fopen(o->path, "wb");
  if (!fp) return -1;
  if (o->is_vlen_str) { const char *s = *(const char *const *)buf; fwrite(s, 1, strlen(s), fp); }
  else if (type == H5T_NATIVE_DOUBLE_g) fwrite(buf, 8, 1, fp);
  else if (type == H5T_NATIVE_INT_g) fwrite(buf, 4, 1, fp);
  else { fclose(fp); return -1; }
  fclose(fp);
  return 0;
}
herr_t H5Aread(hid_t a, hid_t type, void *buf) {
  obj_t *o = get(a, K_ATTR);
  if (!o) return -1;
  const size_t sz = type == H5T_NATIVE_DOUBLE_g ? 8 : (type == H5T_NATIVE_INT_g ? 4 : 0);
  if (!sz) return -1;
  FILE *fp = fopen(o->path, "rb");
  if (!fp) return -1;
  const size_t got = fread(buf, sz, 1, fp);
  fseek(fp, 0, SEEK_END);
  const long len = ftell(fp);
  fclose(fp);
  return (got == 1 && (size_t)len == sz) ? 0 : -1; /* type of the read must match the type written */
}
herr_t H5Aclose(hid_t a) { obj_t *o = get(a, K_ATTR); if (!o) return -1; o->kind = K_FREE; return 0; }

hid_t H5Tcopy(hid_t t) { if (t != H5T_C_S1_g) return -1; return alloc_obj(K_TYPE); }
herr_t H5Tset_size(hid_t t, size_t size) { obj_t *o = get(t, K_TYPE); if (!o) return -1; o->is_vlen_str = size == (size_t)-1; return 0; }
herr_t H5Tclose(hid_t t) { obj_t *o = get(t, K_TYPE); if (!o) return -1; o->kind = K_FREE; return 0; }
hid_t H5Gopen2(hid_t f, const char *name, hid_t gapl) {
  (void)gapl;
  obj_t *fo = get(f, K_FILE);
  if (!fo || strcmp(name, "/") != 0) return -1;
  hid_t id = alloc_obj(K_GROUP);
  if (id >= 0) snprintf(objs[id].path, sizeof(objs[id].path), "%s", get(f, K_FILE)->path);
  return id;
}
herr_t H5Gclose(hid_t g) { obj_t *o = get(g, K_GROUP); if (!o) return -1; o->kind = K_FREE; return 0; }
