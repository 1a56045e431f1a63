// dsp_dynamic_multiple_neighbors.h — drop-in replacement for g-ch/DSP-map include/dsp_dynamic_multiple_neighbors.h: same macros, same `class DSPMap` public surface, the per-frame
// particle loop runs on a B200 through libdspmap_b200.so.  The parameter block below carries the reference's values and
// may be edited (or overridden with -D) exactly like the reference's own (script/set_map_parameters.py rewrites these lines).
#pragma once
#include <string>
/** Parameters for the map **/
#ifndef MAP_LENGTH_VOXEL_NUM
#define MAP_LENGTH_VOXEL_NUM 50
#endif
#ifndef MAP_WIDTH_VOXEL_NUM
#define MAP_WIDTH_VOXEL_NUM 50
#endif
#ifndef MAP_HEIGHT_VOXEL_NUM
#define MAP_HEIGHT_VOXEL_NUM 30
#endif
#ifndef VOXEL_RESOLUTION
#define VOXEL_RESOLUTION 0.2
#endif
#ifndef ANGLE_RESOLUTION
#define ANGLE_RESOLUTION 1
#endif
#ifndef MAX_PARTICLE_NUM_VOXEL
#define MAX_PARTICLE_NUM_VOXEL 30
#endif
#define LIMIT_MOVEMENT_IN_XY_PLANE 1
#ifndef PREDICTION_TIMES
#define PREDICTION_TIMES 6
#define DSPMAP_FUTURE_TIMES {0.05f, 0.2f, 0.5f, 1.f, 1.5f, 2.f}
#endif
#ifndef DSPMAP_HALF_FOV_H
#define DSPMAP_HALF_FOV_H 42
#endif
#ifndef DSPMAP_HALF_FOV_V
#define DSPMAP_HALF_FOV_V 27
#endif
const int half_fov_h = DSPMAP_HALF_FOV_H;
const int half_fov_v = DSPMAP_HALF_FOV_V;
#define DYNAMIC_CLUSTER_MAX_POINT_NUM 200
#define DYNAMIC_CLUSTER_MAX_CENTER_HEIGHT 1.5
static std::string particle_save_folder = ".";
/** END **/
#ifndef PYRAMID_NEIGHBOR_N
#define PYRAMID_NEIGHBOR_N 2
#endif
#define DSPMAP_PYRAMID_NEIGHBOR_N PYRAMID_NEIGHBOR_N
#define DSPMAP_MODEL 0
#define DSPMAP_OCCLUSION_MARGIN ((float)VOXEL_RESOLUTION)
#define DSPMAP_PI_IS_DOUBLE 1  // see dspmap_config.pi_is_double
#define DSPMAP_CSV_SEPARATOR ""  // particle CSV path = particle_save_folder + this + "particles_update_t_..." (mn:335: no separator)
#include "dspmap_dropin.hpp"
