// types/GroundTruthDataRow.h — one 100 Hz row of a preprocessed Jackal log
// (same fields as reference src/types/GroundTruthDataRow.h:3-15; CSV schema src/Trainer.cpp:102-143).
#pragma once

struct GroundTruthDataRow {
    double time = 0.0;            // s
    double vl = 0.0, vr = 0.0;    // commanded left / right wheel speeds, rad/s
    double x = 0.0, y = 0.0, z = 0.0;
    double yaw = 0.0;
    double wz = 0.0;              // yaw rate
    double vx = 0.0, vy = 0.0;    // body-frame velocity at the COM
};
