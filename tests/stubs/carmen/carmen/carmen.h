// tests/stubs/carmen/carmen/carmen.h -- TEST INFRASTRUCTURE: the handful of CARMEN types the reference's carmenwrapper.h names,
// so that gfs-carmen.cpp compiles (never links) against slam-gmapping-learn_b200/host/include in tests/test_consumers_cpu.py
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
typedef void* MSG_INSTANCE;
typedef void* BYTE_ARRAY;
typedef struct { double x, y, theta; } carmen_point_t;
typedef struct { int num_readings; float* range; carmen_point_t laser_pose, robot_pose; double timestamp; } carmen_robot_laser_message;
typedef struct { carmen_point_t truepose, odometrypose; double timestamp; } carmen_simulator_truepos_message;
typedef struct carmen_localize_summary_t* carmen_localize_summary_p;
typedef struct carmen_localize_particle_filter_t* carmen_localize_particle_filter_p;
inline void IPC_listenWait(int) {}
