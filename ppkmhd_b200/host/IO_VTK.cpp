// VTK ImageData (.vti / .pvti) output with the wire format of the reference
// (src/utils/io/IO_VTK.cpp:211-408 single piece; :630-853 + :860-1008 one piece per slab + .pvti).
// Ghost cells are stripped, x runs fastest, the appended section is `_` followed, per variable, by a
// UInt64 byte count and raw little-endian doubles.
#include <cstdint>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <vector>

#include "SolverBase.h"

namespace ppkMHD {
namespace io {

namespace {
std::string padded(int value, int width) {
  std::ostringstream s;
  s << std::setw(width) << std::setfill('0') << value;
  return s.str();
}

struct Piece {  // what differs between the serial and the per-slab file
  const char *vtk_version;
  int ext[6];
};

void write_vti(const std::string &filename, const DataArray3dHost &U, const HydroParams &p, bool ascii, int nbvar,
               const std::map<int, std::string> &names, const Piece &piece) {
  const int gw = p.ghostWidth;
  std::fstream out(filename.c_str(), std::ios_base::out);
  if (ascii) out << "<?xml version=\"1.0\"?>\n";
  out << "<VTKFile type=\"ImageData\" version=\"" << piece.vtk_version
      << "\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n";
  const int *e = piece.ext;
  const bool slab = std::string(piece.vtk_version) == "1.0";
  out << "  <ImageData WholeExtent=\"" << e[0] << " " << e[1] << " " << e[2] << " " << e[3] << " " << e[4] << " " << e[5]
      << "\" " << "Origin=\"" << p.xmin << " " << p.ymin << " " << p.zmin << "\" " << "Spacing=\"" << p.dx << " " << p.dy
      << " " << p.dz << "\">" << (slab ? "" : "") << "\n";
  out << "  <Piece Extent=\"" << e[0] << " " << e[1] << " " << e[2] << " " << e[3] << " " << e[4] << " " << e[5]
      << (slab ? "" : " ") << "\">\n";
  out << "    <PointData>\n";
  out << "    </PointData>\n";
  if (ascii) {
    out << "    <CellData>\n";
    for (int v = 0; v < nbvar; ++v) {
      out << "    <DataArray type=\"Float64\" Name=\"" << names.at(v) << "\" format=\"ascii\" >\n";
      for (int k = gw; k < p.ksize - gw; ++k)
        for (int j = gw; j < p.jsize - gw; ++j)
          for (int i = gw; i < p.isize - gw; ++i) out << U(i, j, k, v) << " ";
      out << "\n    </DataArray>\n";
    }
    out << "    </CellData>\n";
    out << "  </Piece>\n";
    out << "  </ImageData>\n";
    out << "</VTKFile>\n";
    return;
  }
  const uint64_t nbytes = (uint64_t)p.nx * p.ny * p.nz * sizeof(real_t);
  out << "    <CellData>" << std::endl;
  for (int v = 0; v < nbvar; ++v)
    out << "     <DataArray type=\"Float64\" Name=\"" << names.at(v) << "\" format=\"appended\" offset=\""
        << (uint64_t)v * nbytes + (uint64_t)v * sizeof(uint64_t) << "\" />" << std::endl;
  out << "    </CellData>" << std::endl;
  out << "  </Piece>" << std::endl;
  out << "  </ImageData>" << std::endl;
  out << "  <AppendedData encoding=\"raw\">" << std::endl;
  out << "_";
  std::vector<real_t> row((size_t)p.nx);
  for (int v = 0; v < nbvar; ++v) {
    out.write((const char *)&nbytes, sizeof(uint64_t));
    for (int k = gw; k < p.ksize - gw; ++k)
      for (int j = gw; j < p.jsize - gw; ++j) {
        for (int i = 0; i < p.nx; ++i) row[i] = U(i + gw, j, k, v);
        out.write((const char *)row.data(), (std::streamsize)(row.size() * sizeof(real_t)));
      }
  }
  out << "  </AppendedData>" << std::endl;
  out << "</VTKFile>" << std::endl;
}

void write_pvti_header(const std::string &headerFilename, const std::string &outputPrefix, const HydroParams &p, int nbvar,
                       const std::map<int, std::string> &names, int iStep) {
  std::fstream out(headerFilename.c_str(), std::ios_base::out);
  out << "<?xml version=\"1.0\"?>" << std::endl;
  out << "<VTKFile type=\"PImageData\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">" << std::endl;
  out << "  <PImageData WholeExtent=\"" << 0 << " " << p.mx * p.nx << " " << 0 << " " << p.my * p.ny << " " << 0 << " "
      << p.mz * p.nz << "\" GhostLevel=\"0\" " << "Origin=\"" << p.xmin << " " << p.ymin << " " << p.zmin << "\" "
      << "Spacing=\"" << p.dx << " " << p.dy << " " << p.dz << "\">" << std::endl;
  out << "    <PCellData Scalars=\"Scalars_\">" << std::endl;
  for (int v = 0; v < nbvar; ++v) out << "      <PDataArray type=\"Float64\" Name=\"" << names.at(v) << "\"/>" << std::endl;
  out << "    </PCellData>" << std::endl;
  for (int piece = 0; piece < p.nProcs; ++piece) {
    // MPI_Cart coordinates of rank `piece` for dims (mx, my, mz): z fastest (IO_VTK.cpp:975-996 asks the communicator)
    const int cz = piece % p.mz, cy = (piece / p.mz) % p.my, cx = piece / (p.mz * p.my);
    out << " <Piece Extent=\"";
    out << cx * p.nx << " " << cx * p.nx + p.nx << " ";
    out << cy * p.ny << " " << cy * p.ny + p.ny << " ";
    out << cz * p.nz << " " << cz * p.nz + p.nz << " ";
    out << "\" Source=\"" << outputPrefix + "_time" + padded(iStep, 7) + "_mpi" + padded(piece, 5) + ".vti" << "\"/>"
        << std::endl;
  }
  out << "</PImageData>" << std::endl;
  out << "</VTKFile>" << std::endl;
}
}  // namespace

void save_VTK_3D(const DataArray3dHost &Uhost, HydroParams &params, ConfigMap &configMap, int nbvar,
                 const std::map<int, std::string> &variables_names, int iStep, const std::string &debug_name) {
  const std::string dir = configMap.getString("output", "outputDir", "./");
  const std::string prefix = configMap.getString("output", "outputPrefix", "output");
  const bool ascii = configMap.getBool("output", "outputVtkAscii", false);
  const std::string filename = debug_name.empty() ? dir + "/" + prefix + "_" + padded(iStep, 7) + ".vti"
                                                  : dir + "/" + prefix + "_" + debug_name + "_" + padded(iStep, 7) + ".vti";
  Piece piece{"0.1", {0, params.nx, 0, params.ny, 0, params.nz}};
  write_vti(filename, Uhost, params, ascii, nbvar, variables_names, piece);
}

void save_VTK_3D_slab(const DataArray3dHost &Uhost, HydroParams &params, ConfigMap &configMap, int nbvar,
                      const std::map<int, std::string> &variables_names, int iStep, const std::string &debug_name) {
  const std::string dir = configMap.getString("output", "outputDir", "./");
  const std::string prefix = configMap.getString("output", "outputPrefix", "output");
  const bool ascii = configMap.getBool("output", "outputVtkAscii", false);
  const std::string tail = "_time" + padded(iStep, 7) + "_mpi" + padded(params.myRank, 5) + ".vti";
  const std::string filename = debug_name.empty() ? dir + "/" + prefix + tail : dir + "/" + prefix + "_" + debug_name + tail;
  if (params.myRank == 0)
    write_pvti_header(dir + "/" + prefix + "_time" + padded(iStep, 7) + ".pvti", prefix, params, nbvar, variables_names, iStep);
  const int x0 = params.myMpiPos[0] * params.nx, y0 = params.myMpiPos[1] * params.ny, z0 = params.myMpiPos[2] * params.nz;
  Piece piece{"1.0", {x0, x0 + params.nx, y0, y0 + params.ny, z0, z0 + params.nz}};
  write_vti(filename, Uhost, params, ascii, nbvar, variables_names, piece);
}

// save_VTK_2D (src/utils/io/IO_VTK.cpp:24-206): version "1.0" header also in a serial run, z extent "0 0", spacing "dx dy 0",
// a blank before the closing quote of the piece extent
void save_VTK_2D(const DataArray3dHost &U, HydroParams &p, ConfigMap &configMap, int nbvar,
                 const std::map<int, std::string> &names, int iStep, const std::string &debug_name) {
  const std::string dir = configMap.getString("output", "outputDir", "./");
  const std::string prefix = configMap.getString("output", "outputPrefix", "output");
  const bool ascii = configMap.getBool("output", "outputVtkAscii", false);
  const std::string filename = debug_name.empty() ? dir + "/" + prefix + "_" + padded(iStep, 7) + ".vti"
                                                  : dir + "/" + prefix + "_" + debug_name + "_" + padded(iStep, 7) + ".vti";
  const int gw = p.ghostWidth;
  std::fstream out(filename.c_str(), std::ios_base::out);
  if (ascii) out << "<?xml version=\"1.0\"?>\n";
  out << "<VTKFile type=\"ImageData\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n";
  out << "  <ImageData WholeExtent=\"" << 0 << " " << p.nx << " " << 0 << " " << p.ny << " " << 0 << " " << 0 << "\" "
      << "Origin=\"" << p.xmin << " " << p.ymin << " " << 0 << "\" " << "Spacing=\"" << p.dx << " " << p.dy << " " << 0.0
      << "\">\n";
  out << "  <Piece Extent=\"" << 0 << " " << p.nx << " " << 0 << " " << p.ny << " " << 0 << " " << 0 << " " << "\">\n";
  out << "    <PointData>\n";
  out << "    </PointData>\n";
  if (ascii) {
    out << "    <CellData>\n";
    for (int v = 0; v < nbvar; ++v) {
      out << "    <DataArray type=\"Float64\" Name=\"" << names.at(v) << "\" format=\"ascii\" >\n";
      for (int j = gw; j < p.jsize - gw; ++j)
        for (int i = gw; i < p.isize - gw; ++i) out << U(i, j, 0, v) << " ";
      out << "\n    </DataArray>\n";
    }
    out << "    </CellData>\n";
    out << "  </Piece>\n";
    out << "  </ImageData>\n";
    out << "</VTKFile>\n";
    return;
  }
  const uint64_t nbytes = (uint64_t)p.nx * p.ny * sizeof(real_t);
  out << "    <CellData>" << std::endl;
  for (int v = 0; v < nbvar; ++v)
    out << "     <DataArray type=\"Float64\" Name=\"" << names.at(v) << "\" format=\"appended\" offset=\""
        << (uint64_t)v * nbytes + (uint64_t)v * sizeof(uint64_t) << "\" />" << std::endl;
  out << "    </CellData>" << std::endl;
  out << "  </Piece>" << std::endl;
  out << "  </ImageData>" << std::endl;
  out << "  <AppendedData encoding=\"raw\">" << std::endl;
  out << "_";
  for (int v = 0; v < nbvar; ++v) {
    out.write((const char *)&nbytes, sizeof(uint64_t));
    for (int j = gw; j < p.jsize - gw; ++j)
      for (int i = gw; i < p.isize - gw; ++i) {
        const real_t tmp = U(i, j, 0, v);
        out.write((const char *)&tmp, sizeof(real_t));
      }
  }
  out << "  </AppendedData>" << std::endl;
  out << "</VTKFile>" << std::endl;
}

IO_ReadWrite::IO_ReadWrite(HydroParams &params_, ConfigMap &configMap_, std::map<int, std::string> &names)
  : params(params_), configMap(configMap_), variables_names(names) {
  vtk_enabled = configMap.getBool("output", "vtk_enabled", true);    // IO_ReadWrite.cpp:37
  hdf5_enabled = configMap.getBool("output", "hdf5_enabled", false);  // compile-gated in the reference (USE_HDF5)
}

bool IO_ReadWrite::load_data(DataArray3dHost &Uhost, int &iStep, real_t &time, std::string *why) {
  const std::string inputFilename = configMap.getString("run", "restart_filename", "");
  const std::string h5(".h5");
  if (inputFilename.size() < h5.size() || inputFilename.compare(inputFilename.size() - h5.size(), h5.size(), h5) != 0) {
    if (why) *why = "[run] restart_filename must name a .h5 file (the reference's other format, PnetCDF, needs MPI)";
    return false;
  }
  return load_HDF5(Uhost, params, configMap, variables_names, inputFilename, iStep, time, why);
}

void IO_ReadWrite::save_data(DataArray3dHost &Uhost, int iStep, real_t time, const std::string &debug_name) {
  if (vtk_enabled) {
    if (params.dimType == TWO_D) save_VTK_2D(Uhost, params, configMap, params.nbvar, variables_names, iStep, debug_name);
    else if (params.nProcs > 1) save_VTK_3D_slab(Uhost, params, configMap, params.nbvar, variables_names, iStep, debug_name);
    else save_VTK_3D(Uhost, params, configMap, params.nbvar, variables_names, iStep, debug_name);
  }
  if (hdf5_enabled) {  // Save_HDF5 (IO_ReadWrite.cpp:108-125); libhdf5 is looked up at run time
    std::string why;
    if (!save_HDF5(Uhost, params, configMap, variables_names, iStep, time, &why)) {
      if (!hdf5_failed && params.myRank == 0) std::cerr << "HDF5 output skipped: " << why << std::endl;
      hdf5_failed = true;
    }
  }
}

}  // namespace io
}  // namespace ppkMHD
