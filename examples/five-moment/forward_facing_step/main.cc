// Forward-facing step (Woodward & Colella): Mach 3 flow in a channel of length 3 and height 1 with a step of height 0.2
// at x = 0.6.  The GPU-path counterpart of the reference's extension example
// (/root/reference: examples/five-moment/forward_facing_step/main.cc): an extension object populates the triangulation and
// names the six boundaries, the input file does the rest.
//
// Same domain, same `RefinementFactor` entry, same boundary ids (0 left, 1 bottom, 2 step face, 3 step top, 4 right,
// 5 top).  Difference: the reference grafts concentric shells of small cells around the step corner
// (src/grid_generation.cc); this mesh is the uniform 15 x 5 (times RefinementFactor) channel with the step cells removed.
//
//   make -C warpii_b200 bin/forward_facing_step
//   warpii_b200/bin/forward_facing_step examples/five-moment/mach3_step.inp
#include <cmath>
#include <limits>

#include "warpii_cli.hpp"

using namespace warpii_b200;

class ForwardFacingStep : public GridExtension {
   public:
    void declare_geometry_parameters(ParameterFile& prm) override {
        prm.declare_entry("RefinementFactor", "1", ParameterFile::Pattern::Integer(1));
    }

    void populate_triangulation(Triangulation2D& tria, const ParameterFile& prm) override {
        const int refinement_factor = (int)prm.get_integer("RefinementFactor");
        const double Lx = 3.0, Ly = 1.0;
        const int nx = 15, ny = 5;
        const double dx = Lx / nx, dy = Ly / ny;
        // the channel minus the step: cells with x > 3 dx and y < dy
        tria.subdivided_rectangle(nx * refinement_factor, ny * refinement_factor, 0.0, 0.0, Lx, Ly,
                                  [&](double x, double y) { return x > 3 * dx && y < dy; });
        const double tol = std::sqrt(std::numeric_limits<double>::epsilon());
        for (int c = 0; c < (int)tria.cells.size(); c++)
            for (int f = 0; f < 4; f++) {
                const auto mid = tria.face_center(c, f);
                int id = -1;
                if (std::abs(mid[0]) < tol) id = 0;                                        // left
                else if (std::abs(mid[1]) < tol) id = 1;                                   // bottom
                else if (std::abs(mid[0] - 3 * dx) < tol && mid[1] < dy) id = 2;           // step face
                else if (std::abs(mid[1] - dy) < tol && mid[0] > 3 * dx) id = 3;           // step top
                else if (std::abs(mid[0] - Lx) < tol) id = 4;                              // right
                else if (std::abs(mid[1] - Ly) < tol) id = 5;                              // top
                if (id >= 0) tria.boundary_ids[{c, f}] = id;   // ids on interior faces are ignored
            }
    }
};

int main(int argc, char** argv) { return warpii_cli_main(argc, argv, std::make_shared<ForwardFacingStep>()); }
