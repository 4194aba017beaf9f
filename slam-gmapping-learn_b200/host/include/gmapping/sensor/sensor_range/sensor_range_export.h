// sensor_range_export.h -- symbol visibility macro (the reference generates this header with cmake's generate_export_header,
// openslam_gmapping/CMakeLists.txt:181-196); everything is exported from the one shared library here.
#pragma once
#define SENSOR_RANGE_EXPORT __attribute__((visibility("default")))
