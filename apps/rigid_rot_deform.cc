// rigid_rot_deform source.obj reference.obj output.obj [GRID_RESOLUTION=64] [MESH_RESOLUTION=5000] [lambda=1] [symmetry=0]
// (reference src/app/rigid_rot_deform.cc): EdgeLossWithRot rigidity with per-vertex rotations, Deformer::DeformWithRot.
#include "deform_main.h"
int main(int argc, char** argv) { return mo_app::deform_main(argc, argv, "rigid_rot_deform", true); }
