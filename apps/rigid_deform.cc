// rigid_deform source.obj reference.obj output.obj [GRID_RESOLUTION=64] [MESH_RESOLUTION=5000] [lambda=1] [symmetry=0]
// (reference src/app/rigid_deform.cc): EdgeLoss rigidity, Deformer::Deform.
#include "deform_main.h"
int main(int argc, char** argv) { return mo_app::deform_main(argc, argv, "rigid_deform", false); }
