/* <starneig/expert.h> -- expert configuration, Hessenberg part only.
 * Drop-in for reference src/include/starneig/expert.h:60-99 (other stages are out of scope). */
#ifndef STARNEIG_EXPERT_H
#define STARNEIG_EXPERT_H

#include <starneig/configuration.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STARNEIG_HESSENBERG_DEFAULT_TILE_SIZE    -1     /* expert.h:67 */
#define STARNEIG_HESSENBERG_DEFAULT_PANEL_WIDTH  -1     /* expert.h:72 */

/* expert.h:77-92. tile_size is validated (>= 8 or default) exactly as the reference does
 * (src/hessenberg/interface.c:62-72) but has no effect: the GPU works on the dense column-major
 * matrix, there are no tiles. panel_width is the number of columns reduced per panel
 * (default: interface.c:74-78). */
struct starneig_hessenberg_conf {
    int tile_size;
    int panel_width;
};

/* expert.h:99 / interface.c:131-135 */
void starneig_hessenberg_init_conf(struct starneig_hessenberg_conf *conf);

#ifdef __cplusplus
}
#endif

#endif
