// synth_demo.cpp — exercises the drop-in host API end to end without the (private) Jackal logs: writes synthetic
// Train3/CV3/LD3-style CSV files (linear kinematic model of the reference's scripts/gen_fake_data.py), then runs
// the same calls as the reference's evaluate.cpp and train.cpp (trainThreads + train) against them.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <random>
#include <string>
#include <vector>
#include <algorithm>

#include "TestTerrainMaps.h"
#include "Trainer.h"
#include "VehicleSystem.h"

static void writeCsv(const std::string& fn, int rows, unsigned seed) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(0, 1);
    const double B[3][2] = {{0.0472, 0.0438}, {0.0036, -0.0035}, {-0.1744, 0.1748}};
    const double vbar = 2 + 4 * U(rng), dl = 2 * U(rng) - 1, f = 0.05 + 0.4 * U(rng), ph = 6.28 * U(rng);
    double x = 0, y = 0, yaw = 0;
    std::ofstream o(fn);
    o << ",time,vel_left,vel_right,x,y,yaw,wx,wy,wz,vx,vy\n";
    o.precision(12);
    for (int i = 0; i < rows; i++) {
        const double t = 0.01 * i, vl = vbar + dl + std::sin(6.28 * f * t + ph), vr = vbar - dl + std::sin(6.28 * f * t);
        const double vx = B[0][0] * vl + B[0][1] * vr, vy = B[1][0] * vl + B[1][1] * vr, wz = B[2][0] * vl + B[2][1] * vr;
        o << i << "," << t << "," << vl << "," << vr << "," << x << "," << y << "," << yaw << ",0,0," << wz << "," << vx << "," << vy << "\n";
        for (int k = 0; k < 10; k++) {
            x += 1e-3 * (std::cos(yaw) * vx - std::sin(yaw) * vy);
            y += 1e-3 * (std::sin(yaw) * vx + std::cos(yaw) * vy);
            yaw += 1e-3 * wz;
        }
    }
}

// usage: synth_demo [dir] [--gpus G]    (G > 1: trainThreads shards the batch over G GPUs, one all-reduce per step)
int main(int argc, char** argv) {
    std::string dir = "/tmp/avsl_synth/";
    int gpus = 0;
    for (int i = 1; i < argc; i++) {
        if (std::string(argv[i]) == "--gpus" && i + 1 < argc) gpus = std::atoi(argv[++i]);
        else dir = argv[i];
    }
    char buf[256];
    std::snprintf(buf, sizeof(buf), "mkdir -p %s", dir.c_str());
    if (std::system(buf) != 0) return 2;
    for (int i = 1; i <= 17; i++) { std::snprintf(buf, sizeof(buf), "%sTrain3_data%02d.csv", dir.c_str(), i); writeCsv(buf, 1300, 100 + i); }
    std::snprintf(buf, sizeof(buf), "%sLD3_data01.csv", dir.c_str());
    writeCsv(buf, 2500, 7);

    auto map = std::make_shared<const FlatTerrainMap<ADF>>();
    auto factory = std::make_shared<VehicleSystemFactory<ADF>>(map);
    Trainer train(factory, 2);
    if (gpus > 0) train.setNumGpus(gpus);
    std::cout << "GPUs: " << train.numGpus() << "\n";
    train.setDataDir(dir);
    train.setParamFile(dir + "tire.net");
    train.load();
    train.evaluate_ld3();
    train.evaluate_train3();
    for (int it = 0; it < 3; it++) {
        train.trainThreads();  // one full-batch step over 16 files x 6 windows
        train.evaluate_validation_dataset();
    }
    train.save();
    {   // order-independent fingerprint of the trained parameters (compared across --gpus settings and against the Python path)
        std::cout.precision(17);
        double sum = 0, wsum = 0;
        for (size_t i = 0; i < train.m_params.size(); i++) { sum += train.m_params[i]; wsum += (double)(i + 1) * train.m_params[i]; }
        std::cout << "params sum " << sum << " weighted " << wsum << " p[0] " << train.m_params[0] << " p[103] " << train.m_params[103] << "\n";
    }
    {   // regression (round-1 advice): parameters set on the host after a device-side optimiser step must win, and a step taken
        // through one handle must survive another handle re-binding the shared context
        auto a = factory->makeSystem(), b = factory->makeSystem();
        const int np = a->getNumParams();
        System<ADF>::VectorS p0(np), stepped(np), again(np), custom(np), back(np);
        a->getParams(p0);
        std::vector<double> red(np + 2, 0.0);
        for (int i = 0; i < 104; i++) red[i] = 0.5 * ((i % 7) - 3);
        red[np + 1] = 1.0;  // n_kept
        a->applyRmsProp(red.data(), 1e-3, 0.0);
        a->getParams(stepped);
        b->setParams(p0);       // another handle takes the context over ...
        b->getParams(again);
        a->getParams(again);    // ... and the step taken through `a` is still there
        custom = stepped;
        custom[0] += 0.125;
        a->setParams(custom);   // host parameters after a device step: must not be ignored
        a->getParams(back);
        bool ok = stepped[1] != p0[1] && stepped[104] == p0[104];
        for (int i = 0; i < np; i++) ok = ok && again[i] == stepped[i] && back[i] == custom[i];
        // and they are the ones the device evaluates with: a forward evaluation differs from the one with the stepped set
        System<ADF>::VectorS X(23), Xd1(23), Xd2(23);
        a->getDefaultInitialState(X);
        X[17] = X[19] = 4.0; X[18] = X[20] = 3.0; X[21] = 4.0; X[22] = 3.0;
        a->forward(X, Xd1);
        a->setParams(stepped);
        a->forward(X, Xd2);
        double dmax = 0;
        for (int i = 0; i < 21; i++) dmax = std::max(dmax, std::fabs(Xd1[i] - Xd2[i]));
        ok = ok && dmax > 0;
        std::cout << "parameter hand-over check: " << (ok ? "OK" : "FAILED") << " (|dXd| " << dmax << ")\n";
        if (!ok) return 3;
    }
    std::cout << "synth_demo OK\n";
    return 0;
}
