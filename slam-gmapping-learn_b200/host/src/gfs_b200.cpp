// gfs_b200 -- log-driven front end of the B200 filter (the role of the reference's gfs_nogui / GridSlamProcessorThread
// loop, gui/gsp_thread.cpp:305-453): CARMEN log in, optional .gfs trace out, one JSON line with timing on stdout.
//
//   gfs_b200 -filename in.log [-outfilename out.gfs] [-particles N] [-maxUrange m] [-maxrange m] [-sigma s] [-kernelSize k]
//            [-lstep m] [-astep rad] [-iterations n] [-lsigma s] [-ogain g] [-lskip n] [-srr v -srt v -str v -stt v]
//            [-linearUpdate m] [-angularUpdate rad] [-resampleThreshold f] [-xmin m -ymin m -xmax m -ymax m -delta m]
//            [-randseed s] [-scans n] [-no-generateMap] [-verbose]
//            [-lobsGain g] [-generateMap] [-autosize] [-minimumScore s] [-maxMove m] [-regscore s] [-critscore s]
//            [-linearOdometryReliability v] [-angularOdometryReliability v] [-llsamplerange v] [-llsamplestep v]
//            [-lasamplerange v] [-lasamplestep v]                      (the option names of gui/gsp_thread.cpp:97-144)
// Defaults are gfs-carmen's (gfs-carmen/gfs-carmen.cpp:52-84).  -autosize sizes the map like gfs_nogui does
// (gsp_thread.cpp:315-324, 634-642): the extent of the poses in the log (SensorLog::boundingBox) grown by three times the
// maximum range, and the filter starts at the first range reading's pose.
#include <gmapping/gridfastslam/gridslamprocessor.h>
#include <gmapping/log/carmenconfiguration.h>
#include <gmapping/log/sensorlog.h>
#include <gmapping/log/sensorstream.h>
#include <gmapping/utils/stat.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

using namespace GMapping;

int main(int argc, char** argv) {
    std::map<std::string, double> num = {
        {"particles", 30}, {"maxUrange", 80}, {"maxrange", 80}, {"sigma", 0.05}, {"kernelSize", 1}, {"lstep", 0.05}, {"astep", 0.05},
        {"iterations", 5}, {"lsigma", 0.075}, {"ogain", 3}, {"lskip", 0}, {"srr", 0.01}, {"srt", 0.01}, {"str", 0.01}, {"stt", 0.01},
        {"linearUpdate", 1.0}, {"angularUpdate", 0.5}, {"resampleThreshold", 0.5}, {"xmin", -100}, {"ymin", -100}, {"xmax", 100},
        {"ymax", 100}, {"delta", 0.05}, {"randseed", 0}, {"scans", 1e9}, {"minimumScore", 0}, {"period", -1},
        {"maxMove", 1}, {"regscore", 1e4}, {"critscore", 0}, {"linearOdometryReliability", 0}, {"angularOdometryReliability", 0},
        {"llsamplerange", 0.01}, {"llsamplestep", 0.01}, {"lasamplerange", 0.005}, {"lasamplestep", 0.005}};
    std::string filename, outfilename;
    bool generateMap = true, verbose = false, autosize = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-no-generateMap") { generateMap = false; continue; }
        if (a == "-generateMap") { generateMap = true; continue; }
        if (a == "-autosize") { autosize = true; continue; }
        if (a == "-lobsGain") a = "-ogain";  // gfs_nogui's name for the likelihood gain
        if (a == "-verbose") { verbose = true; continue; }
        if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); return 2; }
        if (a == "-filename") { filename = argv[++i]; continue; }
        if (a == "-outfilename") { outfilename = argv[++i]; continue; }
        const std::string key = a.substr(1);
        if (a[0] != '-' || !num.count(key)) { fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
        num[key] = atof(argv[++i]);
    }
    if (filename.empty()) { fprintf(stderr, "usage: gfs_b200 -filename in.log [-outfilename out.gfs] [options]\n"); return 2; }

    std::ifstream cfgStream(filename.c_str());
    if (!cfgStream) { fprintf(stderr, "cannot open %s\n", filename.c_str()); return 1; }
    CarmenConfiguration conf;
    conf.load(cfgStream);
    cfgStream.close();
    SensorMap sensorMap = conf.computeSensorMap();
    std::ifstream logStream(filename.c_str());
    InputSensorStream input(sensorMap, logStream);

    std::ofstream devnull("/dev/null");
    GridSlamProcessor gsp(verbose ? std::cout : devnull);
    gsp.setSensorMap(sensorMap);
    gsp.setMatchingParameters(num["maxUrange"], num["maxrange"], num["sigma"], (int)num["kernelSize"], num["lstep"], num["astep"],
                              (int)num["iterations"], num["lsigma"], num["ogain"], (unsigned int)num["lskip"]);
    gsp.setMotionModelParameters(num["srr"], num["srt"], num["str"], num["stt"]);
    gsp.setUpdateDistances(num["linearUpdate"], num["angularUpdate"], num["resampleThreshold"]);
    gsp.setUpdatePeriod(num["period"]);
    gsp.setgenerateMap(generateMap);
    gsp.setminimumScore(num["minimumScore"]);
    gsp.setmaxMove(num["maxMove"]);
    gsp.setregScore(num["regscore"]);
    gsp.setcritScore(num["critscore"]);
    gsp.m_matcher.setlinearOdometryReliability(num["linearOdometryReliability"]);
    gsp.m_matcher.setangularOdometryReliability(num["angularOdometryReliability"]);
    gsp.setllsamplerange(num["llsamplerange"]);
    gsp.setllsamplestep(num["llsamplestep"]);
    gsp.setlasamplerange(num["lasamplerange"]);
    gsp.setlasamplestep(num["lasamplestep"]);
    OrientedPoint autoStart;
    if (autosize) {
        SensorLog whole(sensorMap);
        std::ifstream all(filename.c_str());
        whole.load(all);
        autoStart = whole.boundingBox(num["xmin"], num["ymin"], num["xmax"], num["ymax"]);
        const double margin = 3 * num["maxrange"];
        num["xmin"] -= margin; num["ymin"] -= margin; num["xmax"] += margin; num["ymax"] += margin;
    }
    if (!outfilename.empty()) gsp.outputStream().open(outfilename.c_str());

    bool inited = false;
    int readN = 0, processedN = 0;
    double msTotal = 0, msMatched = 0, msDevice = 0;
    unsigned long long beamEvals = 0;
    while (input && processedN < (int)num["scans"]) {
        const SensorReading* r = 0;
        input >> r;
        if (!r) continue;
        const RangeReading* rr = dynamic_cast<const RangeReading*>(r);
        if (rr) {
            if (!inited) {
                gsp.init((unsigned int)num["particles"], num["xmin"], num["ymin"], num["xmax"], num["ymax"], num["delta"], autosize ? autoStart : rr->getPose());
                if (num["randseed"] != 0) sampleGaussian(1, (unsigned long)num["randseed"]);
                inited = true;
            }
            const auto t0 = std::chrono::steady_clock::now();
            const bool processed = gsp.processScan(*rr);
            const auto t1 = std::chrono::steady_clock::now();
            readN++;
            if (processed) {
                const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
                msTotal += ms;
                if (processedN > 0) {
                    msMatched += ms;
                    const GridSlamProcessor::DeviceStats d = gsp.deviceStats();
                    msDevice += d.msScanMatch + d.msWeights + d.msRemap + d.msRegister;
                    beamEvals += d.beamEvals;
                }
                processedN++;
            }
        } else if (const OdometryReading* o = dynamic_cast<const OdometryReading*>(r)) {
            if (inited) gsp.processTruePos(*o);
        }
        delete r;
    }
    const int best = inited ? gsp.getBestParticleIndex() : -1;
    const int matched = processedN > 1 ? processedN - 1 : 0;
    printf("{\"impl\": \"b200\", \"particles\": %d, \"read\": %d, \"processed\": %d, \"ms_total\": %.3f, \"ms_per_matched_scan\": %.4f, "
           "\"device_ms_per_matched_scan\": %.4f, \"beam_evals_per_matched_scan\": %.0f, \"best_particle\": %d, \"map\": [%.3f, %.3f, %.3f, %.3f]}\n",
           (int)num["particles"], readN, processedN, msTotal, matched ? msMatched / matched : 0.0, matched ? msDevice / matched : 0.0,
           matched ? (double)beamEvals / matched : 0.0, best, num["xmin"], num["ymin"], num["xmax"], num["ymax"]);
    return 0;
}
