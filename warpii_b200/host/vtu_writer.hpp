// solution_<n>.vtu frames of the FiveMoment application.
//
// The reference writes them through deal.II's DataOut (five_moment.h:245-315): the solution components named
// <species>_density, <species>_{x,y,z}_momentum, <species>_energy, E_field x3, B_field x3, ph_maxwell_gauss_error,
// ph_maxwell_monopole_error, the derived fields of FiveMomentPostprocessor (postprocessor.h:33-62: x/y/z_velocity,
// pressure, specific_entropy = log p - gamma log rho, speed_of_sound) and the cell's "owner" rank, on patches of linear
// sub-cells (write_higher_order_cells = false).  This writer produces the same kind of file without deal.II: a VTK XML
// UnstructuredGrid whose points are the Gauss-Lobatto nodes of every element (discontinuous across elements, as DG data
// are), connected into (Np-1)^dim linear sub-cells per element, data as raw appended Float64.  Differences: the sub-cell
// vertices are the GLL nodes themselves (no re-interpolation to equidistant points), and with more than one species the
// derived fields are written per species with the species name as prefix (the reference's post-processor only accepts
// the single-species layout).
#pragma once
#include <cmath>
#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace warpii_b200 {

struct VtuSpecies {
    std::string name;
};

class VtuWriter {
   public:
    // state[elem][comp][node] and xyz[elem][node][dim] in the same (device) order; nc = 5 * n_species (+ 8 fields)
    static void write(const std::string& path, int dim, int fe_degree, int64_t n_elems, int nc, const std::vector<VtuSpecies>& species,
                      bool fields_enabled, double gas_gamma, int owner_rank, const double* state, const double* xyz) {
        const int Np = fe_degree + 1;
        int NN = 1, nsub = 1;
        for (int d = 0; d < dim; d++) { NN *= Np; nsub *= Np - 1; }
        const int64_t n_points = n_elems * NN, n_cells = n_elems * nsub;
        const int verts = 1 << dim;
        const uint8_t cell_type = dim == 1 ? 3 : (dim == 2 ? 9 : 12);   // VTK_LINE, VTK_QUAD, VTK_HEXAHEDRON

        // geometry
        Array points("Points", "Float64", 3);
        points.f64.resize((size_t)n_points * 3, 0.0);
        for (int64_t i = 0; i < n_points; i++)
            for (int d = 0; d < dim; d++) points.f64[(size_t)i * 3 + d] = xyz[(size_t)i * dim + d];
        Array conn("connectivity", "Int64", 1), offs("offsets", "Int64", 1), types("types", "UInt8", 1);
        conn.i64.reserve((size_t)n_cells * verts);
        offs.i64.reserve((size_t)n_cells);
        types.u8.assign((size_t)n_cells, cell_type);
        for (int64_t e = 0; e < n_elems; e++) {
            const int64_t base = e * NN;
            for (int s = 0; s < nsub; s++) {
                int c[3] = {0, 0, 0}, t = s;
                for (int d = 0; d < dim; d++) { c[d] = t % (Np - 1); t /= Np - 1; }
                auto node = [&](int a, int b, int cc) { return base + (c[0] + a) + Np * ((dim > 1 ? c[1] + b : 0) + Np * (dim > 2 ? c[2] + cc : 0)); };
                if (dim == 1) { conn.i64.push_back(node(0, 0, 0)); conn.i64.push_back(node(1, 0, 0)); }
                else {
                    for (int k = 0; k < (dim == 3 ? 2 : 1); k++) {   // VTK vertex order: counter-clockwise, bottom then top
                        conn.i64.push_back(node(0, 0, k)); conn.i64.push_back(node(1, 0, k));
                        conn.i64.push_back(node(1, 1, k)); conn.i64.push_back(node(0, 1, k));
                    }
                }
                offs.i64.push_back((int64_t)conn.i64.size());
            }
        }
        // point data
        std::vector<Array> pdata;
        auto component = [&](const std::string& name, int comp) {
            Array a(name, "Float64", 1);
            a.f64.resize((size_t)n_points);
            for (int64_t e = 0; e < n_elems; e++)
                for (int j = 0; j < NN; j++) a.f64[(size_t)e * NN + j] = state[((size_t)e * nc + comp) * NN + j];
            pdata.push_back(std::move(a));
        };
        const int n_species = (int)species.size();
        for (int s = 0; s < n_species; s++) {
            const std::string& n = species[s].name;
            component(n + "_density", 5 * s);
            component(n + "_x_momentum", 5 * s + 1);
            component(n + "_y_momentum", 5 * s + 2);
            component(n + "_z_momentum", 5 * s + 3);
            component(n + "_energy", 5 * s + 4);
        }
        if (fields_enabled) {
            const char* names[8] = {"E_field_x", "E_field_y", "E_field_z", "B_field_x", "B_field_y", "B_field_z",
                                    "ph_maxwell_gauss_error", "ph_maxwell_monopole_error"};
            for (int k = 0; k < 8; k++) component(names[k], 5 * n_species + k);
        }
        for (int s = 0; s < n_species; s++) {   // FiveMomentPostprocessor::evaluate_vector_field
            const std::string prefix = n_species > 1 ? species[s].name + "_" : std::string();
            const char* names[6] = {"x_velocity", "y_velocity", "z_velocity", "pressure", "specific_entropy", "speed_of_sound"};
            std::vector<Array> d(6);
            for (int k = 0; k < 6; k++) { d[k] = Array(prefix + names[k], "Float64", 1); d[k].f64.resize((size_t)n_points); }
            for (int64_t e = 0; e < n_elems; e++)
                for (int j = 0; j < NN; j++) {
                    const double* q = state + ((size_t)e * nc + 5 * s) * NN + j;
                    const double rho = q[0], mx = q[NN], my = q[2 * (size_t)NN], mz = q[3 * (size_t)NN], E = q[4 * (size_t)NN];
                    const double p = (gas_gamma - 1) * (E - (mx * mx + my * my + mz * mz) / (2 * rho));   // euler.h:32-44
                    const size_t i = (size_t)e * NN + j;
                    d[0].f64[i] = mx / rho; d[1].f64[i] = my / rho; d[2].f64[i] = mz / rho;
                    d[3].f64[i] = p;
                    d[4].f64[i] = std::log(p) - gas_gamma * std::log(rho);
                    d[5].f64[i] = std::sqrt(gas_gamma * p / rho);
                }
            for (auto& a : d) pdata.push_back(std::move(a));
        }
        Array owner("owner", "Float64", 1);
        owner.f64.assign((size_t)n_cells, (double)owner_rank);

        // ---- file: header with offsets into one raw appended block ---------------------------------------------
        std::ofstream out(path, std::ios::binary);
        if (!out) throw std::runtime_error("cannot write " + path);
        uint64_t offset = 0;
        auto decl = [&](const Array& a) {
            out << "        <DataArray type=\"" << a.type << "\" Name=\"" << a.name << "\" NumberOfComponents=\"" << a.ncomp
                << "\" format=\"appended\" offset=\"" << offset << "\"/>\n";
            offset += sizeof(uint64_t) + a.bytes();
        };
        out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
            << "  <UnstructuredGrid>\n    <Piece NumberOfPoints=\"" << n_points << "\" NumberOfCells=\"" << n_cells << "\">\n";
        out << "      <Points>\n";
        decl(points);
        out << "      </Points>\n      <Cells>\n";
        decl(conn); decl(offs); decl(types);
        out << "      </Cells>\n      <PointData>\n";
        for (const Array& a : pdata) decl(a);
        out << "      </PointData>\n      <CellData>\n";
        decl(owner);
        out << "      </CellData>\n    </Piece>\n  </UnstructuredGrid>\n  <AppendedData encoding=\"raw\">\n_";
        auto dump = [&](const Array& a) {
            const uint64_t n = a.bytes();
            out.write(reinterpret_cast<const char*>(&n), sizeof n);
            if (!a.f64.empty()) out.write(reinterpret_cast<const char*>(a.f64.data()), (std::streamsize)n);
            else if (!a.i64.empty()) out.write(reinterpret_cast<const char*>(a.i64.data()), (std::streamsize)n);
            else if (!a.u8.empty()) out.write(reinterpret_cast<const char*>(a.u8.data()), (std::streamsize)n);
        };
        dump(points); dump(conn); dump(offs); dump(types);
        for (const Array& a : pdata) dump(a);
        dump(owner);
        out << "\n  </AppendedData>\n</VTKFile>\n";
        if (!out) throw std::runtime_error("error while writing " + path);
    }

    // Names of the point-data arrays in file order (solution components, fields, derived fields).
    static std::vector<std::string> point_data_names(const std::vector<VtuSpecies>& species, bool fields_enabled) {
        std::vector<std::string> out;
        const int n_species = (int)species.size();
        for (const VtuSpecies& sp : species)
            for (const char* c : {"_density", "_x_momentum", "_y_momentum", "_z_momentum", "_energy"}) out.push_back(sp.name + c);
        if (fields_enabled)
            for (const char* c : {"E_field_x", "E_field_y", "E_field_z", "B_field_x", "B_field_y", "B_field_z", "ph_maxwell_gauss_error",
                                  "ph_maxwell_monopole_error"})
                out.push_back(c);
        for (const VtuSpecies& sp : species)
            for (const char* c : {"x_velocity", "y_velocity", "z_velocity", "pressure", "specific_entropy", "speed_of_sound"})
                out.push_back((n_species > 1 ? sp.name + "_" : std::string()) + c);
        return out;
    }

    // solution_<n>.pvtu: the index a sharded run's rank 0 writes next to the per-rank pieces (what DataOut's
    // write_vtu_in_parallel / write_pvtu_record provide in the reference, five_moment.h:312-314).
    static void write_pvtu(const std::string& path, const std::vector<VtuSpecies>& species, bool fields_enabled,
                           const std::vector<std::string>& piece_files) {
        std::ofstream out(path);
        if (!out) throw std::runtime_error("cannot write " + path);
        out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
            << "  <PUnstructuredGrid GhostLevel=\"0\">\n    <PPoints>\n      <PDataArray type=\"Float64\" Name=\"Points\" NumberOfComponents=\"3\"/>\n"
            << "    </PPoints>\n    <PPointData>\n";
        for (const std::string& n : point_data_names(species, fields_enabled))
            out << "      <PDataArray type=\"Float64\" Name=\"" << n << "\" NumberOfComponents=\"1\"/>\n";
        out << "    </PPointData>\n    <PCellData>\n      <PDataArray type=\"Float64\" Name=\"owner\" NumberOfComponents=\"1\"/>\n    </PCellData>\n";
        for (const std::string& f : piece_files) out << "    <Piece Source=\"" << f << "\"/>\n";
        out << "  </PUnstructuredGrid>\n</VTKFile>\n";
    }

   private:
    struct Array {
        Array() = default;
        Array(std::string n, std::string t, int nc) : name(std::move(n)), type(std::move(t)), ncomp(nc) {}
        std::string name, type;
        int ncomp = 1;
        std::vector<double> f64;
        std::vector<int64_t> i64;
        std::vector<uint8_t> u8;
        uint64_t bytes() const { return f64.size() * 8 + i64.size() * 8 + u8.size(); }
    };
};

}  // namespace warpii_b200
