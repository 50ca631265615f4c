// FiveMomentGpuApp: the FiveMoment application of WarpII driven from the reference's own input-file format, on the GPU path.
//
// Mirrors FiveMomentWrapper::create_app / FiveMomentApp<dim>::{declare_parameters, create_from_parameters, setup, run}
// (five_moment.cc:13-52, five_moment.h:99-243), Species / SpeciesFunc (species.cc:9-66, species_func.cc:9-51) and the
// HyperRectangle grid description (grid_descriptions.cc:27-49): same entries, defaults, patterns and two-pass parsing.
// Differences, all stated where they occur: only GridType = HyperRectangle is supported; n_dims = 3 is accepted (the
// reference's case 3 is commented out); the VTU frames come from an own writer (vtu_writer.hpp) instead of DataOut.
#pragma once
#include <cstdio>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "dg_solver.hpp"
#include "expression.hpp"
#include "parameter_file.hpp"
#include "vtu_writer.hpp"

namespace warpii_b200 {

// SpeciesFunc<dim> (species_func.h): a 5-component parsed function in primitive [rho, ux, uy, uz, p] or conserved
// variables; value() always returns conserved components.
class SpeciesFunc {
   public:
    static void declare_parameters(ParameterFile& prm, int dim) {
        prm.declare_entry("VariablesType", "Primitive", ParameterFile::Pattern::Selection("Primitive|Conserved"));
        // Functions::ParsedFunction<dim>::declare_parameters(prm, 5)
        prm.declare_entry("Variable names", dim == 1 ? "x,t" : (dim == 2 ? "x,y,t" : "x,y,z,t"));
        prm.declare_entry("Function constants", "");
        prm.declare_entry("Function expression", "0; 0; 0; 0; 0");
    }
    static std::unique_ptr<SpeciesFunc> create_from_parameters(const ParameterFile& prm, int dim, double gas_gamma) {
        std::unique_ptr<SpeciesFunc> f(new SpeciesFunc());
        f->primitive_ = prm.get("VariablesType") == "Primitive";
        f->gas_gamma_ = gas_gamma;
        f->dim_ = dim;
        std::vector<std::string> vars;
        for (const std::string& v : ParameterFile::split(prm.get("Variable names"), ',')) vars.push_back(ParameterFile::trim(v));
        if ((int)vars.size() != dim + 1 && (int)vars.size() != dim)
            throw std::invalid_argument("Variable names: expected " + std::to_string(dim) + " coordinates and the time variable");
        f->time_index_ = (int)vars.size() == dim + 1 ? dim : -1;
        const auto constants = Expression::parse_constants(prm.get("Function constants"));
        const auto comps = Expression::split_components(prm.get("Function expression"));
        if (comps.size() != 5)
            throw std::invalid_argument("Function expression: the number of components (5) is not equal to the number of expressions (" +
                                        std::to_string(comps.size()) + ")");
        for (const std::string& c : comps) f->components_.emplace_back(new Expression(c, vars, constants));
        return f;
    }
    bool time_dependent() const {
        if (time_index_ < 0) return false;
        for (const auto& c : components_)
            if (c->uses_variable(time_index_)) return true;
        return false;
    }
    // conserved state at (x, t): SpeciesFunc::value for all five components (species_func.cc:9-30)
    void conserved(const double* x, double t, double* q5) const {
        double args[4] = {0, 0, 0, 0};
        for (int d = 0; d < dim_; d++) args[d] = x[d];
        if (time_index_ >= 0) args[time_index_] = t;
        double v[5];
        for (int c = 0; c < 5; c++) v[c] = components_[c]->eval(args);
        if (!primitive_) {
            for (int c = 0; c < 5; c++) q5[c] = v[c];
            return;
        }
        const double rho = v[0];
        q5[0] = rho;
        double kinetic_energy = 0.0;
        for (int d = 0; d < 3; d++) {
            q5[d + 1] = rho * v[d + 1];
            kinetic_energy += 0.5 * rho * v[d + 1] * v[d + 1];
        }
        q5[4] = kinetic_energy + v[4] / (gas_gamma_ - 1);
    }

   private:
    SpeciesFunc() = default;
    bool primitive_ = true;
    double gas_gamma_ = 5.0 / 3.0;
    int dim_ = 1, time_index_ = -1;
    std::vector<std::unique_ptr<Expression>> components_;
};

struct SpeciesDescription {
    std::string name = "neutral";
    double charge = 0.0, mass = 1.0;
    std::vector<int32_t> bc_kind;                          // per boundary id
    std::vector<std::shared_ptr<SpeciesFunc>> inflow;      // per boundary id (null unless Inflow)
    std::shared_ptr<SpeciesFunc> initial_condition;
};

class FiveMomentGpuApp {
   public:
    // FiveMomentWrapper::declare_parameters + create_app (five_moment.cc:13-52)
    // ext: the reference's `Warpii::create_from_cli(argc, argv, std::make_shared<MyExtension>())` (warpii.h; used by
    // examples/five-moment/forward_facing_step/main.cc) -- needed when the input says GridType = Extension
    static std::unique_ptr<FiveMomentGpuApp> create_from_input(const std::string& input, int rank = 0, int n_ranks = 1, int device = 0,
                                                               std::shared_ptr<GridExtension> ext = nullptr) {
        using Pat = ParameterFile::Pattern;
        ParameterFile prm;
        prm.declare_entry("WorkDir", "%A__%I");                                          // warpii.cc:139-149
        prm.declare_entry("Application", "FiveMoment", Pat::Selection("FiveMoment|FPETest"));
        prm.parse_input_from_string(input, true);
        if (prm.get("Application") != "FiveMoment") throw std::invalid_argument("Application must be FiveMoment for the GPU path");
        prm.declare_entry("n_dims", "1", Pat::Integer(1, 3));
        prm.declare_entry("n_species", "1", Pat::Integer());
        prm.declare_entry("n_boundaries", "0", Pat::Integer());
        prm.enter_subsection("geometry");                                                // grid.cc:14-21
        prm.declare_entry("GridType", "HyperRectangle", Pat::Selection("HyperRectangle|ForwardFacingStep|Extension"));
        prm.leave_subsection();
        prm.parse_input_from_string(input, true);

        const int dim = (int)prm.get_integer("n_dims");
        const int n_species = (int)prm.get_integer("n_species");
        const int n_boundaries = (int)prm.get_integer("n_boundaries");
        if (n_species < 1) throw std::invalid_argument("n_species must be at least 1");
        if (n_boundaries < 0) throw std::invalid_argument("n_boundaries must not be negative");

        // FiveMomentApp<dim>::declare_parameters (five_moment.h:99-147)
        for (int i = 0; i < n_species; i++) {
            prm.enter_subsection("Species_" + std::to_string(i + 1));
            declare_species(prm, n_boundaries, dim);
            prm.leave_subsection();
        }
        prm.enter_subsection("geometry");
        const std::string grid_type = prm.get("GridType");
        if (grid_type == "Extension") {
            // grid.cc: the extension declares its own entries inside `geometry` and later fills the triangulation
            if (!ext) throw std::invalid_argument("GridType = Extension needs a grid extension (create_from_input(..., ext))");
            if (dim != 2) throw std::invalid_argument("GridType = Extension: extension grids are two-dimensional on the GPU path");
            if (n_ranks != 1) throw std::invalid_argument("GridType = Extension runs on one GPU");
            ext->declare_geometry_parameters(prm);
        } else if (grid_type != "HyperRectangle") {
            throw std::invalid_argument("GridType = " + grid_type + " is not supported by the GPU path (HyperRectangle or Extension)");
        }
        if (grid_type == "HyperRectangle") {   // HyperRectangleDescription<dim>::declare_parameters (grid_descriptions.cc:27-33)
            std::string zeros, ones, nx1;
            for (int d = 0; d < dim; d++) {
                zeros += d ? ", 0" : "0";
                ones += d ? ", 1" : "1";
                nx1 += d ? ", 1" : "1";
            }
            prm.declare_entry("left", zeros);
            prm.declare_entry("right", ones);
            prm.declare_entry("nx", nx1);
            prm.declare_entry("periodic_dimensions", "x,y,z", Pat::MultipleSelection("x|y|z"));
        }
        prm.leave_subsection();
        prm.declare_entry("fe_degree", "2", Pat::Integer(1, 6));
        prm.declare_entry("fields_enabled", "auto", Pat::Selection("true|false|auto"));
        prm.declare_entry("gas_gamma", "1.6666666666667", Pat::Double());
        prm.declare_entry("t_end", "0.0", Pat::Double(0.0));
        prm.declare_entry("write_output", "true", Pat::Bool());
        prm.declare_entry("n_writeout_frames", "10", Pat::Integer(0));
        // Extension of the input format (no counterpart in the reference, whose operator leaves the fields untouched):
        // the two-fluid source terms of north_star kernel 4, see include/warpii_gpu.h, warpii_gpu_set_sources.
        prm.declare_entry("five_moment_sources", "false", Pat::Bool());
        prm.declare_entry("epsilon0", "1.0", Pat::Double(1e-300));
        prm.declare_entry("phm_chi", "0.0", Pat::Double(0.0));
        // ... and the perfectly hyperbolic Maxwell fluxes that evolve the field components (warpii_gpu_set_maxwell): light
        // speed and the magnetic cleaning speed (phm_chi above is the electric one); the fields start from zero
        prm.declare_entry("five_moment_maxwell", "false", Pat::Bool());
        prm.declare_entry("light_speed", "1.0", Pat::Double(1e-300));
        prm.declare_entry("phm_gamma", "0.0", Pat::Double(0.0));
        prm.parse_input_from_string(input, false);

        // FiveMomentApp<dim>::create_from_parameters (five_moment.h:149-198)
        std::unique_ptr<FiveMomentGpuApp> app(new FiveMomentGpuApp());
        app->dim_ = dim;
        app->n_species_ = n_species;
        app->n_boundaries_ = n_boundaries;
        app->gas_gamma_ = prm.get_double("gas_gamma");
        for (int i = 0; i < n_species; i++) {
            prm.enter_subsection("Species_" + std::to_string(i + 1));
            app->species_.push_back(create_species(prm, n_boundaries, dim, app->gas_gamma_));
            prm.leave_subsection();
        }
        prm.enter_subsection("geometry");
        BoxDescription box;
        box.dim = dim;
        Triangulation2D tria;
        if (grid_type == "Extension") {
            ext->populate_triangulation(tria, prm);
            if (tria.cells.empty()) throw std::invalid_argument("GridType = Extension: the extension produced no cells");
        } else {
            const std::vector<double> nx = ParameterFile::to_doubles(prm.get("nx"));
            const std::vector<double> left = ParameterFile::to_doubles(prm.get("left"));
            const std::vector<double> right = ParameterFile::to_doubles(prm.get("right"));
            if ((int)nx.size() != dim || (int)left.size() != dim || (int)right.size() != dim)
                throw std::invalid_argument("geometry: left, right and nx need " + std::to_string(dim) + " entries");
            const std::string periodic = prm.get("periodic_dimensions");
            const char* names = "xyz";
            for (int d = 0; d < dim; d++) {
                if (nx[d] < 1 || nx[d] != std::floor(nx[d])) throw std::invalid_argument("geometry: nx must hold positive integers");
                box.nx[d] = (int)nx[d];
                box.left[d] = left[d];
                box.right[d] = right[d];
                box.periodic[d] = periodic.find(names[d]) != std::string::npos;   // grid_descriptions.cc:60-71
            }
        }
        prm.leave_subsection();
        app->box_ = box;
        app->fe_degree_ = (int)prm.get_integer("fe_degree");
        const std::string fields = prm.get("fields_enabled");
        app->fields_enabled_ = fields == "true" || (fields == "auto" && n_species > 1);
        app->t_end_ = prm.get_double("t_end");
        app->write_output_ = prm.get_bool("write_output");
        app->n_writeout_frames_ = (int)prm.get_integer("n_writeout_frames");
        app->workdir_format_ = prm.get("WorkDir");
        app->sources_ = prm.get_bool("five_moment_sources");
        app->epsilon0_ = prm.get_double("epsilon0");
        app->chi_ = prm.get_double("phm_chi");
        app->maxwell_ = prm.get_bool("five_moment_maxwell");
        app->light_speed_ = prm.get_double("light_speed");
        app->phm_gamma_ = prm.get_double("phm_gamma");
        if ((app->sources_ || app->maxwell_) && !app->fields_enabled_)
            throw std::invalid_argument("five_moment_sources / five_moment_maxwell = true needs the field components (fields_enabled)");
        if (app->maxwell_ && grid_type == "Extension")
            throw std::invalid_argument("five_moment_maxwell = true is implemented on HyperRectangle grids only");
        app->rank_ = rank;
        app->n_ranks_ = n_ranks;

        std::vector<SpeciesBC> bcs(n_species);
        for (int s = 0; s < n_species; s++) {
            const SpeciesDescription& sp = app->species_[s];
            bcs[s].kind = sp.bc_kind;
            bcs[s].inflow.assign(n_boundaries, std::array<double, 5>{{0, 0, 0, 0, 0}});
            bcs[s].inflow_function.resize(n_boundaries);
            bcs[s].time_dependent.assign(n_boundaries, false);
            for (int b = 0; b < n_boundaries; b++) {
                if (!sp.inflow[b]) continue;
                std::shared_ptr<SpeciesFunc> f = sp.inflow[b];
                bcs[s].inflow_function[b] = [f](const double* x, double t, double* q5) { f->conserved(x, t, q5); };
                bcs[s].time_dependent[b] = f->time_dependent();
            }
        }
        if (grid_type == "Extension")
            app->solver_ = std::make_shared<FiveMomentGpuSolver>(GeneralMesh::from_triangulation(tria, app->fe_degree_), n_species,
                                                                app->fields_enabled_, app->gas_gamma_, app->t_end_, n_boundaries, bcs,
                                                                device);
        else
            app->solver_ = std::make_shared<FiveMomentGpuSolver>(box, app->fe_degree_, n_species, app->fields_enabled_, app->gas_gamma_,
                                                                app->t_end_, n_boundaries, bcs, rank, n_ranks, device);
        return app;
    }

    // FiveMomentApp::setup (five_moment.h:221-231): grid + solver reinit, initial condition, frame 0
    void setup() {
        solver_->reinit();
        if (sources_) {
            std::vector<double> qm;
            for (const SpeciesDescription& sp : species_) qm.push_back(sp.charge / sp.mass);
            solver_->get_fluid_flux_operator().set_sources(true, epsilon0_, chi_, qm);
        }
        if (maxwell_) solver_->get_fluid_flux_operator().set_maxwell(true, light_speed_, chi_, phm_gamma_);
        for (int s = 0; s < n_species_; s++) {
            std::shared_ptr<SpeciesFunc> ic = species_[s].initial_condition;
            solver_->project_initial_condition(s, [ic](const double* x, double* q5) { ic->conserved(x, 0.0, q5); }, false);
        }
        setup_done_ = true;
        output_results(0, 0.0);
    }

    // FiveMomentApp::run (five_moment.h:233-243)
    void run() {
        if (!setup_done_) setup();
        const double writeout_interval = t_end_ / n_writeout_frames_;
        auto writeout = [&](double t) { output_results((unsigned)std::lround(t / writeout_interval), t); };
        // skip the zeroth writeout because setup() already did it
        TimestepCallback writeout_callback(writeout_interval, writeout, false);
        solver_->solve(writeout_callback);
    }

    // frame_callback is the hook for library users; with write_output every frame also goes to
    // <output_dir>/solution_<frame>.vtu, like the reference's output_results.
    void set_frame_callback(std::function<void(unsigned frame, double t)> cb) { frame_callback_ = std::move(cb); }
    void set_output_dir(const std::string& dir) { output_dir_ = dir; }

    // remove_file_extension (utilities.cc:67-88; UtilitiesTests.RemoveFileExtensionTest): the file name without its
    // directories and without its last extension
    static std::string remove_file_extension(const std::string& filename) {
        const size_t slash = filename.find_last_of('/');
        const std::string file_part = slash == std::string::npos ? filename : filename.substr(slash + 1);
        const size_t dot = file_part.find_last_of('.');
        return dot == std::string::npos ? file_part : file_part.substr(0, dot);
    }

    // format_workdir (warpii.cc:205-219): %A -> application name, %I -> input name without extension ("STDIN" for stdin)
    std::string format_workdir(const std::string& input_name) const {
        std::string out = workdir_format_;
        auto replace_all = [&](const std::string& what, const std::string& with) {
            for (size_t pos = 0; (pos = out.find(what, pos)) != std::string::npos; pos += with.size()) out.replace(pos, what.size(), with);
        };
        replace_all("%A", "FiveMoment");
        replace_all("%I", input_name);
        return out;
    }

    // sharded runs: attach the NCCL communicator (after setup(), before run()); id from warpii_gpu_nccl_unique_id on rank 0
    void attach_comm(const char id[WARPII_GPU_NCCL_ID_BYTES]) { solver_->attach_comm(id); }

    FiveMomentGpuSolver& get_solver() { return *solver_; }
    std::shared_ptr<FiveMomentGpuSolver> solver_ptr() const { return solver_; }
    GpuSolutionVec& get_solution() { return solver_->get_solution(); }
    const std::vector<SpeciesDescription>& species() const { return species_; }
    const BoxDescription& box() const { return box_; }
    int n_dims() const { return dim_; }
    int n_species() const { return n_species_; }
    int n_boundaries() const { return n_boundaries_; }
    int fe_degree() const { return fe_degree_; }
    bool fields_enabled() const { return fields_enabled_; }
    double gas_gamma() const { return gas_gamma_; }
    double t_end() const { return t_end_; }
    bool write_output() const { return write_output_; }
    bool sources_enabled() const { return sources_; }
    double epsilon0() const { return epsilon0_; }
    int n_writeout_frames() const { return n_writeout_frames_; }
    unsigned frames_written() const { return frames_written_; }

   private:
    FiveMomentGpuApp() = default;

    // Species<dim>::declare_parameters (species.cc:9-32)
    static void declare_species(ParameterFile& prm, int n_boundaries, int dim) {
        using Pat = ParameterFile::Pattern;
        prm.declare_entry("name", "neutral", Pat::Selection("neutral|ion|electron"));
        prm.declare_entry("charge", "0.0", Pat::Double());
        prm.declare_entry("mass", "1.0", Pat::Double(0.0));
        prm.enter_subsection("BoundaryConditions");
        for (int i = 0; i < n_boundaries; i++) {
            prm.declare_entry(std::to_string(i), "Wall", Pat::Selection("Wall|Outflow|Inflow"));
            prm.enter_subsection(std::to_string(i) + "_Inflow");
            SpeciesFunc::declare_parameters(prm, dim);
            prm.leave_subsection();
        }
        prm.leave_subsection();
        prm.enter_subsection("InitialCondition");
        SpeciesFunc::declare_parameters(prm, dim);
        prm.leave_subsection();
    }
    // Species<dim>::create_from_parameters (species.cc:34-66)
    static SpeciesDescription create_species(ParameterFile& prm, int n_boundaries, int dim, double gas_gamma) {
        SpeciesDescription sp;
        sp.name = prm.get("name");
        sp.charge = prm.get_double("charge");
        sp.mass = prm.get_double("mass");
        sp.bc_kind.assign(n_boundaries, WARPII_BC_WALL);
        sp.inflow.resize(n_boundaries);
        prm.enter_subsection("BoundaryConditions");
        for (int i = 0; i < n_boundaries; i++) {
            const std::string type = prm.get(std::to_string(i));
            if (type == "Wall") sp.bc_kind[i] = WARPII_BC_WALL;
            else if (type == "Outflow") sp.bc_kind[i] = WARPII_BC_OUTFLOW;
            else {
                sp.bc_kind[i] = WARPII_BC_INFLOW;
                prm.enter_subsection(std::to_string(i) + "_Inflow");
                sp.inflow[i] = SpeciesFunc::create_from_parameters(prm, dim, gas_gamma);
                prm.leave_subsection();
            }
        }
        prm.leave_subsection();
        prm.enter_subsection("InitialCondition");
        sp.initial_condition = SpeciesFunc::create_from_parameters(prm, dim, gas_gamma);
        prm.leave_subsection();
        return sp;
    }

    // FiveMomentApp::output_results (five_moment.h:245-315): the frame callback fires and, with write_output, the state is
    // downloaded (the only device -> host traffic of a run) and written as solution_<n>.vtu (vtu_writer.hpp).
    void output_results(unsigned frame, double t) {
        frames_written_++;
        if (frame_callback_) frame_callback_(frame, t);
        if (!write_output_ || output_dir_.empty()) return;
        std::vector<double> host((size_t)solver_->context()->n_dofs());
        solver_->get_solution().download(host.data());
        if (node_xyz_.empty()) node_xyz_ = solver_->node_coords();
        char name[64];
        if (n_ranks_ > 1) std::snprintf(name, sizeof name, "solution_%03u.rank%d.vtu", frame, rank_);
        else std::snprintf(name, sizeof name, "solution_%03u.vtu", frame);
        std::vector<VtuSpecies> names;
        for (const SpeciesDescription& sp : species_) names.push_back({sp.name});
        VtuWriter::write(output_dir_ + "/" + name, dim_, fe_degree_, solver_->n_local_elems(), solver_->n_components(), names,
                         fields_enabled_, gas_gamma_, rank_, host.data(), node_xyz_.data());
        if (n_ranks_ > 1 && rank_ == 0) {
            std::vector<std::string> pieces;
            for (int r = 0; r < n_ranks_; r++) {
                char piece[64];
                std::snprintf(piece, sizeof piece, "solution_%03u.rank%d.vtu", frame, r);
                pieces.push_back(piece);
            }
            char index_name[64];
            std::snprintf(index_name, sizeof index_name, "/solution_%03u.pvtu", frame);
            VtuWriter::write_pvtu(output_dir_ + index_name, names, fields_enabled_, pieces);
        }
        std::ofstream index(output_dir_ + (n_ranks_ > 1 ? "/frames.rank" + std::to_string(rank_) + ".txt" : std::string("/frames.txt")),
                            frame == 0 ? std::ios::trunc : std::ios::app);
        index.precision(17);
        index << frame << " " << t << " " << name << "\n";
    }

    int dim_ = 1, n_species_ = 1, n_boundaries_ = 0, fe_degree_ = 2, n_writeout_frames_ = 10, rank_ = 0, n_ranks_ = 1;
    bool fields_enabled_ = false, write_output_ = true, setup_done_ = false, sources_ = false, maxwell_ = false;
    double gas_gamma_ = 5.0 / 3.0, t_end_ = 0.0, epsilon0_ = 1.0, chi_ = 0.0, light_speed_ = 1.0, phm_gamma_ = 0.0;
    BoxDescription box_;
    std::vector<SpeciesDescription> species_;
    std::shared_ptr<FiveMomentGpuSolver> solver_;
    std::function<void(unsigned, double)> frame_callback_;
    std::vector<double> node_xyz_;
    std::string output_dir_, workdir_format_ = "%A__%I";
    unsigned frames_written_ = 0;
};

}  // namespace warpii_b200
