/* <starneig/starneig.h> -- umbrella header (reference: src/include/starneig/starneig.h.in). */
#ifndef STARNEIG_STARNEIG_H
#define STARNEIG_STARNEIG_H
#include <starneig/configuration.h>
#include <starneig/error.h>
#include <starneig/node.h>
#include <starneig/expert.h>
#include <starneig/sep_sm.h>
#endif
