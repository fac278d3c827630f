// flatdd_gpu — command line front end with the flags and output contract of the reference's
// apps/FlatDD.cpp (reference apps/FlatDD.cpp:20-146): --file --t/-t --fuse --no_cache --beta --thresh
// --DDSIM_convert --pv --ps, the same stdout markers, the same JSON "statistics" keys and the same
// ../../log/results/{time,state}/<name>_FlatDD.txt files.  The array phase runs on the GPU through
// the C-ABI (include/flatdd_b200.h).  Additions: --gpu D (device), --fuse 3 / 4 / 5 (GPU cost greedy in program order / dense-block fusion / DD-multiply dependency graph),
// --bin FILE (final state as raw little-endian fp64: real array then imag array), --trace FILE
// (also record the boundary traffic), --trace-only (record the boundary traffic WITHOUT a device: the host DD phase,
// the switch rule and the fusion pass run, every flat table that would cross the C-ABI goes to the --trace file; with
// --world N the schedule carries the half-shard exchanges of an N-shard state.  This is how the inputs of bench.py
// are made where no GPU exists), --quiet.
// Multi-GPU: start one process per GPU with --world N --rank r --rendezvous FILE (rank 0 writes the
// NCCL unique id there, the others wait for it); every rank runs the same driver on the same circuit,
// holds the shard with top index bits r and writes it to <--bin>.rank<r>.
#include "cxxopts.hpp"
#include "nlohmann/json.hpp"
#include "reference_binding.hpp"

#include <chrono>
#include <cmath>
#include <fstream>
#include <algorithm>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <thread>

int main(int argc, char** argv) {
    cxxopts::Options options("flatdd_gpu", "FlatDD with a B200 array phase");
    // clang-format off
    options.add_options()
        ("h,help", "produce help message")
        ("pv", "save the state vector")
        ("t", "num of threads (kept for CLI compatibility; sizes the reference cost model)", cxxopts::value<unsigned int>()->default_value("16"))
        ("ps", "print simulation stats")
        ("file", "simulate a quantum circuit given by file", cxxopts::value<std::string>())
        ("fuse", "0 off, 1 reference greedy, 2 DATE'19 op count, 3 GPU cost greedy, 4 dense-block fusion for the tile-resident kernel (fastest), 5 dependency-graph fusion by DD multiplication (round 1)", cxxopts::value<unsigned int>()->default_value("0"))
        ("no_cache", "no cache optimization (accepted; the GPU path has no DMAV cache)")
        ("beta", "EMA parameter (parsed and ignored like the reference: beta stays 0.9)", cxxopts::value<double>()->default_value("0.9"))
        ("thresh", "set the threshold", cxxopts::value<double>()->default_value("2"))
        ("DDSIM_convert", "ddsim conversion (accepted; one conversion kernel serves both)")
        ("gpu", "CUDA device", cxxopts::value<int>()->default_value("0"))
        ("bin", "write the final state as raw fp64 (re[], im[])", cxxopts::value<std::string>())
        ("trace", "record the boundary traffic to this file", cxxopts::value<std::string>())
        ("trace-only", "no device: only record the boundary traffic (needs --trace; with --world N the N-shard schedule)")
        ("world", "number of shards / processes (power of two)", cxxopts::value<int>()->default_value("1"))
        ("rank", "rank of this process", cxxopts::value<int>()->default_value("0"))
        ("rendezvous", "file through which rank 0 hands out the NCCL unique id", cxxopts::value<std::string>()->default_value(""))
        ("exchange", "0 = peer-memory kernel, 1 = NCCL send/recv", cxxopts::value<int>()->default_value("0"))
        ("shots", "sample this many measurement outcomes of all qubits on the device (the reference skips measurements)", cxxopts::value<unsigned long>()->default_value("0"))
        ("seed", "seed of the sampler", cxxopts::value<unsigned long>()->default_value("0"))
        ("time-gates", "time every DMAVM launch with CUDA events (synchronises per launch) and report achieved HBM GB/s")
        ("quiet", "no progress output");
    // clang-format on
    auto vm = options.parse(argc, argv);
    if (vm.count("help") > 0 || vm.count("file") == 0) {
        std::cout << options.help();
        return vm.count("help") > 0 ? 0 : 1;
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    const std::string fname = vm["file"].as<std::string>();
    auto circuit = std::make_unique<qc::QuantumComputation>(fname);
    const int nQubits = static_cast<int>(circuit->getNqubits());

    const int world = vm["world"].as<int>();
    const int rank = vm["rank"].as<int>();
    const bool traceOnly = vm.count("trace-only") > 0;
    if (traceOnly && vm.count("trace") == 0) {
        std::cerr << "flatdd_gpu: --trace-only needs --trace FILE\n";
        return 1;
    }
    std::unique_ptr<fddb200::GpuArrayBackend> gpu;
    try {
        if (traceOnly) {
            // no device is touched; nothing is computed
        } else if (world > 1) {
            const std::string rendezvous = vm["rendezvous"].as<std::string>();
            if (rendezvous.empty()) throw std::runtime_error("--world needs --rendezvous FILE");
            char id[128];
            if (rank == 0) {
                fddb200::fddCheck(fdd_comm_unique_id(id), "fdd_comm_unique_id");
                std::ofstream tmp(rendezvous + ".tmp", std::ios::binary);
                tmp.write(id, sizeof id);
                tmp.close();
                std::rename((rendezvous + ".tmp").c_str(), rendezvous.c_str());
            } else {
                for (int tries = 0;; ++tries) {
                    std::ifstream in(rendezvous, std::ios::binary);
                    if (in.read(id, sizeof id)) break;
                    if (tries > 6000) throw std::runtime_error("rendezvous file did not appear: " + rendezvous);
                    std::this_thread::sleep_for(std::chrono::milliseconds(10));
                }
            }
            const int device = vm.count("gpu") > 0 && vm["gpu"].count() > 0 ? vm["gpu"].as<int>() : rank;
            gpu = std::make_unique<fddb200::GpuArrayBackend>(nQubits, device, rank, world, id, vm["exchange"].as<int>());
        } else {
            gpu = std::make_unique<fddb200::GpuArrayBackend>(nQubits, vm["gpu"].as<int>());
        }
    } catch (const std::exception& e) {
        std::cerr << "flatdd_gpu: " << e.what() << "\n";
        return 3; // no CPU fallback
    }
    std::unique_ptr<fddb200::TraceRecorder> recorder;
    std::unique_ptr<fddb200::TeeBackend> tee;
    fddb200::ArrayBackend* backend = gpu.get();
    if (vm.count("trace") > 0) {
        recorder = std::make_unique<fddb200::TraceRecorder>(vm["trace"].as<std::string>(), nQubits);
        if (traceOnly) {
            if (world > 1) recorder->setWorldSize(world);
            backend = recorder.get();
        } else {
            tee = std::make_unique<fddb200::TeeBackend>(std::vector<fddb200::ArrayBackend*>{gpu.get(), recorder.get()});
            backend = tee.get();
        }
    }
    if (vm.count("time-gates") > 0 && gpu) gpu->setTiming(true);
    fddb200::RefGpuSwitchSimulator sim(std::move(circuit), backend);
    sim.threshold = vm["thresh"].as<double>();
    const auto nThread = vm["t"].as<unsigned int>();
    sim.n_thread_exp = static_cast<unsigned int>(std::log2(nThread));
    sim.fuse = vm["fuse"].as<unsigned int>();
    sim.enable_cache = vm.count("no_cache") == 0;
    sim.ddsim_convert = vm.count("DDSIM_convert") > 0;
    sim.verbose = vm.count("quiet") == 0 && rank == 0;
    sim.worldSize = world;
    sim.timePerGate = vm.count("time-gates") > 0;

    sim.simulate();
    const auto t2 = std::chrono::high_resolution_clock::now();
    const std::chrono::duration<float> durationSimulation = t2 - t1;
    std::cout << "Simulation finished" << std::endl;

    if (traceOnly) {
        if (!sim.switched) sim.getVectorFromDD(); // what --pv / getVector would convert (apps/FlatDD.cpp:89-93)
        recorder->close();
        nlohmann::json traced;
        traced["trace"] = {{"file", vm["trace"].as<std::string>()}, {"records", recorder->records()}, {"n_qubits", nQubits},
                           {"n_ops", sim.getNumberOfOps()}, {"switched", sim.switched}, {"switched_at_op", sim.switchedAtOp},
                           {"unitary_ops", sim.unitaryOps}, {"array_phase_ops", sim.arrayPhaseOps}, {"launches", sim.launches},
                           {"world", world}, {"exchanges", sim.exchanges}, {"fuse", sim.fuse}, {"gate_merging_s", sim.gateMergingTime},
                           {"simulation_time", durationSimulation.count()}};
        std::cout << std::setw(2) << traced << std::endl;
        return 0;
    }
    if (vm.count("pv") > 0 || vm.count("bin") > 0) {
        double* re = nullptr;
        double* im = nullptr;
        sim.getVector(re, im); // converts the DD on the device if the switch never fired
        std::size_t dim = std::size_t{1} << sim.getNumberOfQubits();
        for (int w = world; w > 1; w >>= 1) dim >>= 1; // a shard
        if (vm.count("pv") > 0 && world == 1) {
            std::ofstream out("../../log/results/state/" + sim.getName() + "_FlatDD.txt");
            if (out.is_open()) {
                for (std::size_t q = 0; q < dim; ++q) out << re[q] << " " << im[q] << std::endl;
                std::cout << "Data saved to file." << std::endl;
            } else {
                std::cerr << "Failed to open the file." << std::endl;
            }
        }
        if (vm.count("bin") > 0) {
            std::ofstream out(vm["bin"].as<std::string>() + (world > 1 ? ".rank" + std::to_string(rank) : std::string()), std::ios::binary);
            out.write(reinterpret_cast<const char*>(re), static_cast<std::streamsize>(dim * sizeof(double)));
            out.write(reinterpret_cast<const char*>(im), static_cast<std::streamsize>(dim * sizeof(double)));
        }
    }
    if (recorder) recorder->close();

    // measurement sampling on the device: most frequent outcomes as {bitstring: count}
    nlohmann::json sampled;
    const auto shots = vm["shots"].as<unsigned long>();
    if (shots > 0 && world == 1) {
        double* re = nullptr;
        double* im = nullptr;
        if (!sim.switched) sim.getVector(re, im); // make sure the state is on the device
        std::vector<uint64_t> outcomes(shots);
        fddb200::fddCheck(fdd_sample(gpu->ctx(), shots, vm["seed"].as<unsigned long>(), outcomes.data()), "fdd_sample");
        std::map<uint64_t, std::size_t> histogram;
        for (uint64_t o : outcomes) ++histogram[o];
        std::vector<std::pair<std::size_t, uint64_t>> byCount;
        for (const auto& kv : histogram) byCount.emplace_back(kv.second, kv.first);
        std::sort(byCount.rbegin(), byCount.rend());
        for (std::size_t k = 0; k < byCount.size() && k < 16; ++k) {
            std::string bits(static_cast<std::size_t>(nQubits), '0');
            for (int q = 0; q < nQubits; ++q) {
                if ((byCount[k].second >> q) & 1U) bits[static_cast<std::size_t>(nQubits - 1 - q)] = '1';
            }
            sampled[bits] = byCount[k].first;
        }
    }

    nlohmann::json outputObj;
    outputObj["statistics"] = {{"simulation_time", durationSimulation.count()},
                               {"benchmark", sim.getName()},
                               {"n_qubits", +sim.getNumberOfQubits()},
                               {"applied_gates", sim.getNumberOfOps()},
                               {"DD->Array conversion", sim.getSwitchTime()},
                               {"number of threads", nThread},
                               // additions
                               {"switched", sim.switched},
                               {"switched_at_op", sim.switchedAtOp},
                               {"unitary_gates", sim.unitaryOps},
                               {"array_phase_gates", sim.arrayPhaseOps},
                               {"array_phase_launches", sim.launches},
                               {"array_phase_time", sim.arrayPhaseTime},
                               {"gate_merging_time", sim.gateMergingTime},
                               {"gates_per_sec_array_phase", sim.arrayPhaseTime > 0 ? static_cast<double>(sim.arrayPhaseOps) / sim.arrayPhaseTime : 0.0},
                               {"gpu_kernel_launches", fdd_launch_count(gpu->ctx())},
                               {"dmavm_kernel_ms_total", sim.kernelMsTotal},
                               {"dmavm_hbm_gbs", sim.kernelMsTotal > 0 ? 32.0 * std::ldexp(1.0, nQubits) / (world > 1 ? world : 1) * static_cast<double>(sim.launches) / (sim.kernelMsTotal * 1e-3) / 1e9 : 0.0},
                               {"world", world},
                               {"rank", rank},
                               {"exchanges", sim.exchanges}};
    if (!sampled.empty()) outputObj["samples_top16"] = sampled;
    if (rank != 0) return 0; // rank 0 reports
    std::ofstream timeFile("../../log/results/time/" + sim.getName() + "_FlatDD.txt");
    if (timeFile.is_open()) {
        for (const auto& t : sim.getTimeRecord1()) timeFile << t << std::endl;
        timeFile << "Switch Overhead:" << sim.getSwitchTime() << std::endl;
        for (const auto& t : sim.getTimeRecord2()) timeFile << t << std::endl;
        std::cout << "Time data saved to file." << std::endl;
    } else {
        std::cerr << "Failed to open the time file." << std::endl;
    }
    std::cout << std::setw(2) << outputObj << std::endl;
    return 0;
}
