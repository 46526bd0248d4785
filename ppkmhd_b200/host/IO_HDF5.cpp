// HDF5 output / restart input + the XDMF wrapper of the reference
//   writeXdmfForHdf5Wrapper        src/utils/io/IO_HDF5.cpp:16-378   (plain text, needs no library)
//   Save_HDF5<d>::save             src/utils/io/IO_HDF5.h:73-526     (one dataset per variable + 7 root attributes)
//   Load_HDF5<d>::load             src/utils/io/IO_HDF5.h:1537-2153  (restart: datasets -> interior, "time step", "total time")
// The reference links libhdf5 at build time (USE_HDF5). This build has no HDF5 headers or library to link against, so the
// handful of C entry points used are resolved with dlopen / dlsym at run time (HDF5 >= 1.10: 64-bit hid_t); when no
// libhdf5 is installed every HDF5 call reports "unavailable" and the callers say so (the XDMF text is written regardless,
// like the reference does from main.cpp:163-170).
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <vector>

#include "SolverBase.h"

namespace ppkMHD {
namespace io {

namespace {
std::string padded7(int value) {
  std::ostringstream s;
  s << std::setw(7) << std::setfill('0') << value;
  return s.str();
}

// the variables a file holds, in the order the reference writes them (IO_HDF5.h:421-443, IO_HDF5.cpp:223-360)
std::vector<int> file_variables(const HydroParams &p) {
  std::vector<int> v = {ID, IP, IU, IV};
  if (p.mhdEnabled) { v.push_back(IW); v.push_back(IA); v.push_back(IB); v.push_back(IC); }
  else if (p.dimType == THREE_D) v.push_back(IW);
  return v;
}

// ---- libhdf5 through dlopen -------------------------------------------------------------------------------------
typedef int64_t hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;
struct Hdf5Api {
  void *lib = nullptr;
  bool ok = false;
  std::string why;
  herr_t (*H5open)();
  herr_t (*H5get_libversion)(unsigned *, unsigned *, unsigned *);
  hid_t (*H5Fcreate)(const char *, unsigned, hid_t, hid_t);
  hid_t (*H5Fopen)(const char *, unsigned, hid_t);
  herr_t (*H5Fflush)(hid_t, int);
  herr_t (*H5Fclose)(hid_t);
  hid_t (*H5Screate_simple)(int, const hsize_t *, const hsize_t *);
  hid_t (*H5Screate)(int);
  herr_t (*H5Sselect_hyperslab)(hid_t, int, const hsize_t *, const hsize_t *, const hsize_t *, const hsize_t *);
  herr_t (*H5Sclose)(hid_t);
  hid_t (*H5Pcreate)(hid_t);
  herr_t (*H5Pset_chunk)(hid_t, int, const hsize_t *);
  herr_t (*H5Pset_shuffle)(hid_t);
  herr_t (*H5Pset_deflate)(hid_t, unsigned);
  herr_t (*H5Pclose)(hid_t);
  hid_t (*H5Dcreate2)(hid_t, const char *, hid_t, hid_t, hid_t, hid_t, hid_t);
  hid_t (*H5Dopen2)(hid_t, const char *, hid_t);
  herr_t (*H5Dwrite)(hid_t, hid_t, hid_t, hid_t, hid_t, const void *);
  herr_t (*H5Dread)(hid_t, hid_t, hid_t, hid_t, hid_t, void *);
  herr_t (*H5Dclose)(hid_t);
  hid_t (*H5Acreate2)(hid_t, const char *, hid_t, hid_t, hid_t, hid_t);
  hid_t (*H5Aopen)(hid_t, const char *, hid_t);
  herr_t (*H5Awrite)(hid_t, hid_t, const void *);
  herr_t (*H5Aread)(hid_t, hid_t, void *);
  herr_t (*H5Aclose)(hid_t);
  hid_t (*H5Tcopy)(hid_t);
  herr_t (*H5Tset_size)(hid_t, size_t);
  herr_t (*H5Tclose)(hid_t);
  hid_t (*H5Gopen2)(hid_t, const char *, hid_t);
  herr_t (*H5Gclose)(hid_t);
  hid_t NATIVE_DOUBLE = -1, NATIVE_INT = -1, C_S1 = -1, DATASET_CREATE = -1;
};
// values of the public headers (H5Fpublic.h, H5Spublic.h, H5Tpublic.h): stable across 1.10 - 1.14
constexpr unsigned kH5F_ACC_RDONLY = 0u, kH5F_ACC_TRUNC = 2u;
constexpr hid_t kH5P_DEFAULT = 0;
constexpr int kH5S_SCALAR = 0, kH5S_SELECT_SET = 0, kH5F_SCOPE_LOCAL = 0;
constexpr size_t kH5T_VARIABLE = (size_t)-1;

Hdf5Api &hdf5() {
  static Hdf5Api api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  const char *names[] = {getenv("PPK_HDF5_LIB"), "libhdf5.so", "libhdf5_serial.so", "libhdf5.so.310", "libhdf5.so.200",
                         "libhdf5.so.103", "libhdf5_serial.so.103", "libhdf5_serial.so.200"};
  for (const char *nm : names) {
    if (!nm || !*nm) continue;
    api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) {
    api.why = "no libhdf5.so on this machine (set PPK_HDF5_LIB to its path)";
    return api;
  }
  bool all = true;
#define SYM(f)                                         \
  do {                                                 \
    *(void **)(&api.f) = dlsym(api.lib, #f);           \
    if (!api.f) { all = false; api.why = "libhdf5 lacks " #f; } \
  } while (0)
  SYM(H5open); SYM(H5get_libversion); SYM(H5Fcreate); SYM(H5Fopen); SYM(H5Fflush); SYM(H5Fclose); SYM(H5Screate_simple);
  SYM(H5Screate); SYM(H5Sselect_hyperslab); SYM(H5Sclose); SYM(H5Pcreate); SYM(H5Pset_chunk); SYM(H5Pset_shuffle);
  SYM(H5Pset_deflate); SYM(H5Pclose); SYM(H5Dcreate2); SYM(H5Dopen2); SYM(H5Dwrite); SYM(H5Dread); SYM(H5Dclose);
  SYM(H5Acreate2); SYM(H5Aopen); SYM(H5Awrite); SYM(H5Aread); SYM(H5Aclose); SYM(H5Tcopy); SYM(H5Tset_size); SYM(H5Tclose);
  SYM(H5Gopen2); SYM(H5Gclose);
#undef SYM
  if (!all) return api;
  unsigned maj = 0, min = 0, rel = 0;
  if (api.H5open() < 0 || api.H5get_libversion(&maj, &min, &rel) < 0 || maj != 1 || min < 10) {
    api.why = "libhdf5 older than 1.10 (32-bit hid_t) is not supported";
    return api;
  }
  // H5T_NATIVE_DOUBLE & co are macros over global ids that H5open() fills
  auto gid = [&](const char *sym) {
    hid_t *p = (hid_t *)dlsym(api.lib, sym);
    return p ? *p : (hid_t)-1;
  };
  api.NATIVE_DOUBLE = gid("H5T_NATIVE_DOUBLE_g");
  api.NATIVE_INT = gid("H5T_NATIVE_INT_g");
  api.C_S1 = gid("H5T_C_S1_g");
  api.DATASET_CREATE = gid("H5P_CLS_DATASET_CREATE_ID_g");
  if (api.NATIVE_DOUBLE < 0 || api.NATIVE_INT < 0 || api.C_S1 < 0 || api.DATASET_CREATE < 0) {
    api.why = "libhdf5 does not export the predefined type / property-class ids";
    return api;
  }
  api.ok = true;
  return api;
}

void scalar_attr(Hdf5Api &h, hid_t file, const char *name, hid_t type, const void *value) {
  const hid_t ds = h.H5Screate(kH5S_SCALAR);
  const hid_t at = h.H5Acreate2(file, name, type, ds, kH5P_DEFAULT, kH5P_DEFAULT);
  h.H5Awrite(at, type, value);
  h.H5Sclose(ds);
  h.H5Aclose(at);
}
}  // namespace

bool hdf5_available(std::string *why) {
  Hdf5Api &h = hdf5();
  if (why) *why = h.why;
  return h.ok;
}

// writeXdmfForHdf5Wrapper (src/utils/io/IO_HDF5.cpp:16-378): the light-data file that points ParaView at the .h5 files.
// Written in the current directory like the reference (the name carries no outputDir, :128-136).
void writeXdmfForHdf5Wrapper(HydroParams &params, ConfigMap &configMap, const std::map<int, std::string> &variables_names,
                             int totalNumberOfSteps, bool singleStep) {
  const bool ghostIncluded = configMap.getBool("output", "ghostIncluded", false);
  const int gw2 = ghostIncluded ? 2 * params.ghostWidth : 0;
  // a decomposed run re-assembles the pieces into one file in the reference (parallel HDF5): global sizes
  const int nxg = params.nx * params.mx + gw2, nyg = params.ny * params.my + gw2, nzg = params.nz * params.mz + gw2;
  const bool two_d = params.dimType == TWO_D;
  const std::string outputPrefix = configMap.getString("output", "outputPrefix", "output");
  std::string xdmfFilename = outputPrefix + ".xmf";
  if (singleStep) xdmfFilename = outputPrefix + "_" + padded7(totalNumberOfSteps) + ".xmf";
  std::fstream x(xdmfFilename.c_str(), std::ios_base::out);
  x << "<?xml version=\"1.0\" ?>" << std::endl;
  x << "<!DOCTYPE Xdmf SYSTEM \"Xdmf.dtd\" []>" << std::endl;
  x << "<Xdmf xmlns:xi=\"http://www.w3.org/2003/XInclude\" Version=\"2.2\">" << std::endl;
  x << "  <Domain>" << std::endl;
  x << "    <Grid Name=\"TimeSeries\" GridType=\"Collection\" CollectionType=\"Temporal\">" << std::endl;
  const int startStep = singleStep ? totalNumberOfSteps : 0;
  const int stopStep = singleStep ? totalNumberOfSteps + 1 : totalNumberOfSteps;  // (inclusive, as in the reference: :160-163)
  std::ostringstream dims;
  if (two_d) dims << nyg << " " << nxg;
  else dims << nzg << " " << nyg << " " << nxg;
  for (int iStep = startStep; iStep <= stopStep; ++iStep) {
    const std::string baseName = outputPrefix + "_" + padded7(iStep);
    x << "    <Grid Name=\"" << baseName << "\" GridType=\"Uniform\">" << std::endl;
    x << "    <Time Value=\"" << iStep << "\" />" << std::endl;
    x << "      <Topology TopologyType=\"" << (two_d ? "2DCoRectMesh" : "3DCoRectMesh") << "\" NumberOfElements=\"" << dims.str()
      << "\"/>" << std::endl;
    x << "    <Geometry Type=\"" << (two_d ? "ORIGIN_DXDY" : "ORIGIN_DXDYDZ") << "\">" << std::endl;
    for (const char *what : {"Origin", "Spacing"}) {
      x << "    <DataStructure" << std::endl;
      x << "       Name=\"" << what << "\"" << std::endl;
      x << "       DataType=\"Double\"" << std::endl;
      x << "       Dimensions=\"" << (two_d ? 2 : 3) << "\"" << std::endl;
      x << "       Format=\"XML\">" << std::endl;
      const char *v = what[0] == 'O' ? "0" : "1";
      x << "       " << v << " " << v << (two_d ? "" : std::string(" ") + v) << std::endl;
      x << "    </DataStructure>" << std::endl;
    }
    x << "    </Geometry>" << std::endl;
    for (int var : file_variables(params)) {
      x << "      <Attribute Center=\"Node\" Name=\"" << variables_names.at(var) << "\">" << std::endl;
      x << "        <DataStructure" << std::endl;
      x << "           DataType=\"Double\"" << std::endl;
      x << "           Dimensions=\"" << dims.str() << "\"" << std::endl;
      x << "           Format=\"HDF\">" << std::endl;
      x << "           " << baseName << ".h5:/" << variables_names.at(var) << "" << std::endl;
      x << "        </DataStructure>" << std::endl;
      x << "      </Attribute>" << std::endl;
    }
    x << "   </Grid>" << std::endl;
  }
  x << "   </Grid>" << std::endl;
  x << " </Domain>" << std::endl;
  x << "</Xdmf>" << std::endl;
}

// Save_HDF5<d>::save (src/utils/io/IO_HDF5.h:186-526), undecomposed run. Uhost is (isize, jsize, ksize | 1, nbvar), x fastest.
// Returns false (with `why`) when libhdf5 is unavailable or the file cannot be created.
bool save_HDF5(const DataArray3dHost &Uhost, HydroParams &params, ConfigMap &configMap, const std::map<int, std::string> &variables_names,
               int iStep, real_t totalTime, std::string *why) {
  Hdf5Api &h = hdf5();
  if (!h.ok) {
    if (why) *why = h.why;
    return false;
  }
  if (params.nProcs > 1) {
    if (why) *why = "a decomposed run writes one re-assembled file through parallel HDF5 (MPI-IO) in the reference: not available";
    return false;
  }
  const bool two_d = params.dimType == TWO_D;
  const int nx = params.nx, ny = params.ny, nz = params.nz, gw = params.ghostWidth;
  const int isize = params.isize, jsize = params.jsize, ksize = two_d ? 1 : params.ksize;
  const bool ghostIncluded = configMap.getBool("output", "ghostIncluded", false);
  const std::string outputDir = configMap.getString("output", "outputDir", "./");
  const std::string outputPrefix = configMap.getString("output", "outputPrefix", "output");
  const std::string full = outputDir + "/" + outputPrefix + "_" + padded7(iStep) + ".h5";
  const hid_t file = h.H5Fcreate(full.c_str(), kH5F_ACC_TRUNC, kH5P_DEFAULT, kH5P_DEFAULT);
  if (file < 0) {
    if (why) *why = "H5Fcreate failed for " + full;
    return false;
  }
  const int rank = two_d ? 2 : 3;
  const int g2 = ghostIncluded ? 2 * gw : 0;
  // slowest dimension first (row-major files over the x-fastest arrays, :279-301)
  const hsize_t dims_memory[3] = {two_d ? (hsize_t)jsize : (hsize_t)ksize, two_d ? (hsize_t)isize : (hsize_t)jsize, (hsize_t)isize};
  const hsize_t dims_file[3] = {two_d ? (hsize_t)(ny + g2) : (hsize_t)(nz + g2), two_d ? (hsize_t)(nx + g2) : (hsize_t)(ny + g2),
                                (hsize_t)(nx + g2)};
  const hid_t space_mem = h.H5Screate_simple(rank, dims_memory, nullptr);
  const hid_t space_file = h.H5Screate_simple(rank, dims_file, nullptr);
  const hsize_t off = ghostIncluded ? 0 : (hsize_t)gw;
  const hsize_t start[3] = {off, off, off}, stride[3] = {1, 1, 1}, block[3] = {1, 1, 1};
  h.H5Sselect_hyperslab(space_mem, kH5S_SELECT_SET, start, stride, dims_file, block);
  int level = (int)configMap.getInteger("output", "outputHdf5CompressionLevel", 0);
  if (level < 0 || level > 9) {
    std::cerr << "Invalid value for compression level; must be an integer between 0 and 9 !!!" << std::endl;
    std::cerr << "compression level is then set to default value 0; i.e. no compression !!" << std::endl;
    level = 0;
  }
  const hid_t plist = h.H5Pcreate(h.DATASET_CREATE);
  const hsize_t chunk[3] = {two_d ? (hsize_t)ny : (hsize_t)nz, two_d ? (hsize_t)nx : (hsize_t)ny, (hsize_t)nx};
  h.H5Pset_chunk(plist, rank, chunk);
  h.H5Pset_shuffle(plist);
  h.H5Pset_deflate(plist, (unsigned)level);
  bool ok = true;
  const size_t per_var = (size_t)isize * jsize * ksize;
  for (int var : file_variables(params)) {
    const std::string name = "/" + variables_names.at(var);
    const hid_t ds = h.H5Dcreate2(file, name.c_str(), h.NATIVE_DOUBLE, space_file, kH5P_DEFAULT, plist, kH5P_DEFAULT);
    ok = ok && ds >= 0 && h.H5Dwrite(ds, h.NATIVE_DOUBLE, space_mem, space_file, kH5P_DEFAULT, Uhost.data() + per_var * var) >= 0;
    if (ds >= 0) h.H5Dclose(ds);
  }
  const double timeValue = (double)totalTime;
  const int ghost_flag = ghostIncluded ? 1 : 0;
  scalar_attr(h, file, "time step", h.NATIVE_INT, &iStep);
  scalar_attr(h, file, "total time", h.NATIVE_DOUBLE, &timeValue);
  scalar_attr(h, file, "nx", h.NATIVE_INT, &nx);
  scalar_attr(h, file, "ny", h.NATIVE_INT, &ny);
  scalar_attr(h, file, "nz", h.NATIVE_INT, &nz);
  scalar_attr(h, file, "ghost zone included", h.NATIVE_INT, &ghost_flag);
  {  // "creation date": a variable-length string attribute on the root group (:493-512)
    time_t now = time(nullptr);
    char buf[64];
    strftime(buf, sizeof(buf), "%Y-%m-%d %H:%M:%S", localtime(&now));
    const char *date = buf;
    const hsize_t one[1] = {1};
    const hid_t type = h.H5Tcopy(h.C_S1);
    h.H5Tset_size(type, kH5T_VARIABLE);
    const hid_t root = h.H5Gopen2(file, "/", kH5P_DEFAULT);
    const hid_t sp = h.H5Screate_simple(1, one, nullptr);
    const hid_t at = h.H5Acreate2(root, "creation date", type, sp, kH5P_DEFAULT, kH5P_DEFAULT);
    h.H5Awrite(at, type, &date);
    h.H5Aclose(at);
    h.H5Gclose(root);
    h.H5Tclose(type);
    h.H5Sclose(sp);
  }
  h.H5Pclose(plist);
  h.H5Sclose(space_mem);
  h.H5Sclose(space_file);
  h.H5Fflush(file, kH5F_SCOPE_LOCAL);
  h.H5Fclose(file);
  if (!ok && why) *why = "H5Dcreate2 / H5Dwrite failed for " + full;
  return ok;
}

// Load_HDF5<d>::load (src/utils/io/IO_HDF5.h:1537-2153) for a restart at the same resolution: fills the interior of Uhost
// (the whole array when the file carries its ghost zones), returns the file's "time step" and "total time".
bool load_HDF5(DataArray3dHost &Uhost, HydroParams &params, ConfigMap &configMap, const std::map<int, std::string> &variables_names,
               const std::string &filename, int &iStep, real_t &totalTime, std::string *why) {
  Hdf5Api &h = hdf5();
  if (!h.ok) {
    if (why) *why = h.why;
    return false;
  }
  if (configMap.getBool("run", "restart_upscale", false)) {
    if (why) *why = "restart_upscale (reading a half-resolution file) is not supported";
    return false;
  }
  const hid_t file = h.H5Fopen(filename.c_str(), kH5F_ACC_RDONLY, kH5P_DEFAULT);
  if (file < 0) {
    if (why) *why = "cannot open restart file " + filename;
    return false;
  }
  const bool two_d = params.dimType == TWO_D;
  const int gw = params.ghostWidth;
  const int isize = params.isize, jsize = params.jsize, ksize = two_d ? 1 : params.ksize;
  auto read_attr = [&](const char *name, hid_t type, void *dst) {
    const hid_t at = h.H5Aopen(file, name, kH5P_DEFAULT);
    if (at < 0) return false;
    const bool good = h.H5Aread(at, type, dst) >= 0;
    h.H5Aclose(at);
    return good;
  };
  int fnx = 0, fny = 0, fnz = 0, ghosts = 0, step = 0;
  double t = 0.0;
  bool ok = read_attr("nx", h.NATIVE_INT, &fnx) && read_attr("ny", h.NATIVE_INT, &fny) && read_attr("nz", h.NATIVE_INT, &fnz) &&
            read_attr("ghost zone included", h.NATIVE_INT, &ghosts) && read_attr("time step", h.NATIVE_INT, &step) &&
            read_attr("total time", h.NATIVE_DOUBLE, &t);
  if (ok && (fnx != params.nx || fny != params.ny || (!two_d && fnz != params.nz))) {
    if (why) *why = "restart file has another resolution than [mesh] nx, ny, nz";
    ok = false;
  }
  if (ok) {
    const int rank = two_d ? 2 : 3;
    const int g2 = ghosts ? 2 * gw : 0;
    const hsize_t dims_memory[3] = {two_d ? (hsize_t)jsize : (hsize_t)ksize, two_d ? (hsize_t)isize : (hsize_t)jsize, (hsize_t)isize};
    const hsize_t dims_file[3] = {two_d ? (hsize_t)(fny + g2) : (hsize_t)(fnz + g2), two_d ? (hsize_t)(fnx + g2) : (hsize_t)(fny + g2),
                                  (hsize_t)(fnx + g2)};
    const hid_t space_mem = h.H5Screate_simple(rank, dims_memory, nullptr);
    const hid_t space_file = h.H5Screate_simple(rank, dims_file, nullptr);
    const hsize_t off = ghosts ? 0 : (hsize_t)gw;
    const hsize_t start[3] = {off, off, off}, stride[3] = {1, 1, 1}, block[3] = {1, 1, 1};
    h.H5Sselect_hyperslab(space_mem, kH5S_SELECT_SET, start, stride, dims_file, block);
    const size_t per_var = (size_t)isize * jsize * ksize;
    for (int var : file_variables(params)) {
      const std::string name = "/" + variables_names.at(var);
      const hid_t ds = h.H5Dopen2(file, name.c_str(), kH5P_DEFAULT);
      ok = ok && ds >= 0 && h.H5Dread(ds, h.NATIVE_DOUBLE, space_mem, space_file, kH5P_DEFAULT, Uhost.data() + per_var * var) >= 0;
      if (ds >= 0) h.H5Dclose(ds);
    }
    h.H5Sclose(space_mem);
    h.H5Sclose(space_file);
    if (!ok && why) *why = "a dataset of " + filename + " could not be read";
  } else if (why && why->empty()) {
    *why = "restart file " + filename + " lacks the attributes Save_HDF5 writes";
  }
  h.H5Fclose(file);
  if (ok) {
    iStep = step;
    totalTime = (real_t)t;
  }
  return ok;
}

}  // namespace io
}  // namespace ppkMHD
