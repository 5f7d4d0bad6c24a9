// iS3D_b200.e: same command line as the reference's iS3D.e (reference src/cpp/Main.cpp:4-12) -- run from a
// directory laid out like the reference repository root.
#include <cstdio>

#include "is3d_host.hpp"

int main(int argc, char *argv[])
{
  (void)argc; (void)argv;
  is3dhost::IS3D particlization;
  particlization.run_particlization(1);   // 1 = read the freezeout surface from input/surface.dat
  printf("\nFinished particlization\n");
  return 0;
}
