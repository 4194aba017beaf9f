#include <gmapping/gridfastslam/gridslamprocessor.h>
#include <gmapping/log/carmenconfiguration.h>
#include <gmapping/log/sensorstream.h>
#include <gmapping/utils/stat.h>
#include <fstream>
#include <iostream>
using namespace GMapping;
int main(int argc, char** argv) {
    std::ifstream cfg(argv[1]); CarmenConfiguration conf; conf.load(cfg); cfg.close();
    SensorMap sm = conf.computeSensorMap();
    std::ifstream is(argv[1]); InputSensorStream input(sm, is);
    std::ofstream devnull("/dev/null");
    GridSlamProcessor* gsp = new GridSlamProcessor(devnull);
    gsp->setSensorMap(sm);
    gsp->setMatchingParameters(80, 80, 0.05, 1, 0.05, 0.05, 5, 0.075, 3, 0);
    gsp->setMotionModelParameters(0.1, 0.2, 0.1, 0.2);
    gsp->setUpdateDistances(0.4, 0.2, 0.5);
    gsp->setgenerateMap(true);
    bool inited = false; int n = 0; GridSlamProcessor* copy = 0;
    while (input && n < 14) {
        const SensorReading* r = 0; input >> r; if (!r) continue;
        const RangeReading* rr = dynamic_cast<const RangeReading*>(r);
        if (!rr) { delete r; continue; }
        if (!inited) { gsp->init(8, -100, -100, 100, 100, 0.05, rr->getPose()); sampleGaussian(1, 7); inited = true; }
        gsp->processScan(*rr);
        if (copy) copy->processScan(*rr);
        n++;
        if (n == 6) copy = gsp->clone();
        if (n == 10) { GridSlamProcessor::TNodeVector t = gsp->getTrajectories(); gsp->integrateScanSequence(t[0]); for (size_t i = 0; i < t.size(); i++) delete t[i]; }
        delete r;
    }
    // adaptParticles: a third filter that resamples at every reading and goes 8 -> 5 -> 11 particles
    {
        std::ifstream is2(argv[1]); InputSensorStream input2(sm, is2);
        GridSlamProcessor g3(devnull);
        g3.setSensorMap(sm);
        g3.setMatchingParameters(80, 80, 0.05, 1, 0.05, 0.05, 5, 0.075, 3, 0);
        g3.setMotionModelParameters(0.1, 0.2, 0.1, 0.2);
        g3.setUpdateDistances(0.0, 0.0, 2.0);
        g3.setgenerateMap(true);
        g3.setmaxParticles(12);
        static const int adapt[8] = {0, 0, 0, 5, 5, 11, 0, 0};
        bool up = false; int k = 0;
        while (input2 && k < 8) {
            const SensorReading* r = 0; input2 >> r; if (!r) continue;
            const RangeReading* rr = dynamic_cast<const RangeReading*>(r);
            if (!rr) { delete r; continue; }
            if (!up) { g3.init(8, -100, -100, 100, 100, 0.05, rr->getPose()); up = true; }
            g3.processScan(*rr, adapt[k]);
            k++;
            delete r;
        }
        GridSlamProcessor::TNodeVector t = g3.getTrajectories();
        std::cout << "adapt particles " << g3.getParticles().size() << " trajectories " << t.size() << std::endl;
        for (size_t i = 0; i < t.size(); i++) delete t[i];
        if (g3.getParticles().size() != 11) return 3;
    }
    // read a map through the mirror, copy a particle like the ROS node
    GridSlamProcessor::Particle best = gsp->getParticles()[gsp->getBestParticleIndex()];
    double occ = best.map.cell(best.map.world2map(Point(best.pose.x, best.pose.y)));
    std::cout << "occ " << occ << " neff " << gsp->getneff() << " / " << copy->getneff() << std::endl;
    delete gsp; delete copy;
    return 0;
}
