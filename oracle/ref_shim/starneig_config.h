/* Stand-in for the CMake-generated starneig_config.h (src/CMakeLists.txt configure_file).
 * Test infrastructure only: lets the reference's src/hessenberg/cpu.c compile from where it
 * lies under /root/reference, with MPI/CUDA/events/sanity checks switched off. */
#ifndef STARNEIG_CONFIG_H
#define STARNEIG_CONFIG_H
#define STARNEIG_ENABLE_MESSAGES
#endif
