// aslp-nnet-train-server -- rank 0 of the asynchronous modes: holds the central model, serves the workers in arrival order
// until all of them have sent kMsgFinished, then writes the model.  Same command line and flags as
// src/aslp-parallelbin/aslp-nnet-train-server.cc:24-110 (--server-type=easgd|asgd|masgd, --alpha, --sync-period,
// --masgd-momentum).  Launched as rank 0 of the same torchrun / ASLP_NCCL_ID_FILE group as the worker mains
// (--worker-type=easgd|asgd); the any-source message channel is loopback TCP on ASLP_CTRL_PORT (default MASTER_PORT + 1).
// BatchNorm statistics are reduced at the end only when the net has any (the worker mains do the same), so that the
// collective stays matched.
#include <memory>
#include "nnet-nnet.h"
#include "parallel-async.h"
#include "parse-options.h"

int main(int argc, char* argv[]) {
  using namespace kaldi;
  using namespace kaldi::aslp_nnet;
  try {
    const char* usage =
        "Parameter server for training, it can adapt all kinds of wokers,"
        "eg framewise, sequential and stream training\n"
        "Usage:  aslp-nnet-train-server [options] <model-in> <model-out>\n"
        "e.g.: \n"
        " aslp-nnet-train-server nnet.init nnet.out\n";
    ParseOptions po(usage);
    bool binary = true;
    po.Register("binary", &binary, "Write output in binary mode");
    std::string use_gpu = "yes";
    po.Register("use-gpu", &use_gpu, "yes|no|optional, only has effect if compiled with CUDA");
    std::string server_type = "easgd";
    po.Register("server-type", &server_type, "Server type(easgd | asgd)");
    float alpha = 0.5f;
    po.Register("alpha", &alpha, "Moving rate alpha for easgd server");
    int32 sync_period = 1000;
    po.Register("sync-period", &sync_period, "Synchronization period for ASGD");
    int32 gpu_id = -1;
    po.Register("gpu-id", &gpu_id, "selected gpu id, if negative then select automaticly");
    float masgd_momentum = 0.9f;
    po.Register("masgd-momentum", &masgd_momentum, "momentum for masgd");
    po.Read(argc, argv);
    if (po.NumArgs() != 2) { po.PrintUsage(); return 1; }
    const std::string model_filename = po.GetArg(1), target_model_filename = po.GetArg(2);
    if (use_gpu == "no") KALDI_ERR << "--use-gpu=no: this build has no CPU path";
    if (gpu_id >= 0) ASLP_OK(aslp_set_device(gpu_id));
    else if (const char* lr = std::getenv("LOCAL_RANK")) ASLP_OK(aslp_set_device(std::atoi(lr)));

    Nnet nnet;
    nnet.Read(model_filename);
    WorkerBootstrap boot;
    if (boot.rank != 0) KALDI_ERR << "the server is rank 0 of its group (MpiNode::MainNode), got rank " << boot.rank;
    std::unique_ptr<IServer> server;
    if (server_type == "easgd") server.reset(new EasgdServer(boot.id, boot.nranks, alpha));
    else if (server_type == "asgd") server.reset(new AsgdServer(boot.id, boot.nranks, alpha, sync_period, -1.0f));
    else if (server_type == "masgd") server.reset(new AsgdServer(boot.id, boot.nranks, 1.0f, sync_period, masgd_momentum));
    else KALDI_ERR << "Unsupported server type: " << server_type;
    std::vector<std::pair<BaseFloat*, int>> params;
    nnet.GetGpuParams(&params);
    server->InitParam(params);
    KALDI_LOG << "Mpi cluster info total " << server->NumNodes() << " server rank " << server->Rank();
    server->Run();          // until every worker has finished
    {                       // aslp-nnet-train-server.cc:93-97: the collective the worker mains enter after Stop()
      std::vector<double*> acc_params;
      std::vector<std::pair<double*, int>> data_params;
      nnet.GetAccStats(&acc_params, &data_params);
      server->ReduceAccStat(acc_params, data_params);
    }
    nnet.Write(target_model_filename, binary);
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what();
    return -1;
  }
}
