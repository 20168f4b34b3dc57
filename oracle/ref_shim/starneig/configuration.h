/* Stand-in for the generated <starneig/configuration.h> (src/include/starneig/configuration.h.in). */
#ifndef STARNEIG_CONFIGURATION_H
#define STARNEIG_CONFIGURATION_H
#define STARNEIG_VERSION_MAJOR 0
#define STARNEIG_VERSION_MINOR 2
#define STARNEIG_VERSION_PATCH 0
#endif
