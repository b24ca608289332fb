#!/bin/bash
# One gpurun call of round 2: GPU tests, bench lines, microbenchmarks, live graph timeline, ncu launch list + full capture.
# usage: tools/gpu_call.sh <tag> [steps...]   (steps: tests bench umma eager graph ncu)
tag=$1; shift
steps="${@:-tests bench umma eager graph ncu}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
for s in $steps; do
  case $s in
    guard) # a changed kernel is validated first; when its parity tests fail the remaining steps run on the previous build of the library
      timeout 300 python -m pytest tests -m gpu -q -rf -p no:cacheprovider --timeout 90 --timeout-method=thread -k "$P2I_GUARDSEL" > gpurun_out/${tag}_pytest_guard.txt 2>&1; grc=$?
      echo "guard rc=$grc" >> gpurun_out/${tag}_pytest_guard.txt; tail -8 gpurun_out/${tag}_pytest_guard.txt
      if [ $grc -ne 0 ]; then export P2I_LIB_PATH=$PWD/p2i-gan-benchmark_b200/p2igan_b200/libp2i_sm100a_prev.so; echo "GUARD FAILED: using $P2I_LIB_PATH"; fi;;
    guardb) # second changed kernel family, with a run-time switch back to the previous kernels
      timeout 300 python -m pytest tests -m gpu -q -rf -p no:cacheprovider --timeout 90 --timeout-method=thread -k "$P2I_GUARDSEL_B" > gpurun_out/${tag}_pytest_guardb.txt 2>&1; grc=$?
      echo "guardb rc=$grc" >> gpurun_out/${tag}_pytest_guardb.txt; tail -8 gpurun_out/${tag}_pytest_guardb.txt
      if [ $grc -ne 0 ]; then export P2I_LIB_PATH=$PWD/p2i-gan-benchmark_b200/p2igan_b200/libp2i_sm100a_prev.so; echo "GUARD B FAILED: using $P2I_LIB_PATH"; fi;;
    upsel) # pick the channels-per-thread variant of upmod_bwd_hi_kernel by measurement; the remaining steps run with it
      timeout 200 python tools/bench_upmod.py gpurun_out/${tag}_upmod_best.txt > gpurun_out/${tag}_upmod_ab.txt 2>&1; echo "upsel rc=$?"; cat gpurun_out/${tag}_upmod_ab.txt | tail -6
      if [ -s gpurun_out/${tag}_upmod_best.txt ]; then export P2I_UPMOD_CPT=$(cat gpurun_out/${tag}_upmod_best.txt); echo "P2I_UPMOD_CPT=$P2I_UPMOD_CPT"; fi;;
    tests) timeout 600 python -m pytest tests -m gpu -q -rf -p no:cacheprovider --timeout 120 --timeout-method=thread > gpurun_out/${tag}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.txt; tail -5 gpurun_out/${tag}_pytest.txt;;
    smoke) timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.txt; tail -3 gpurun_out/${tag}_smoke.txt;;
    bench) timeout 900 python bench.py > gpurun_out/${tag}_bench_train.json 2> gpurun_out/${tag}_bench_train.err; echo "bench rc=$?"; head -c 600 gpurun_out/${tag}_bench_train.json;;
    bench20) timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_train20.json 2> gpurun_out/${tag}_bench_train20.err; echo "bench20 rc=$?"; head -c 600 gpurun_out/${tag}_bench_train20.json;;
    prof) P2I_LIB_PATH=$PWD/p2i-gan-benchmark_b200/p2igan_b200/libp2i_sm100a_prof.so timeout 300 python tools/halo_prof.py > gpurun_out/${tag}_halo_prof.txt 2>&1; echo "prof rc=$?"; tail -4 gpurun_out/${tag}_halo_prof.txt;;
    profdbg) for d in 1 3; do P2I_HALO_DBG=$d P2I_LIB_PATH=$PWD/p2i-gan-benchmark_b200/p2igan_b200/libp2i_sm100a_prof.so timeout 300 python tools/halo_prof.py 0 1 > gpurun_out/${tag}_halo_prof_dbg$d.txt 2>&1; echo "profdbg$d rc=$?"; done;;
    pdl0) P2I_PDL=0 timeout 600 python bench.py --steps 100 --no-cpu --no-extras > gpurun_out/${tag}_bench_pdl0.json 2> gpurun_out/${tag}_bench_pdl0.err; echo "pdl0 rc=$?"; head -c 300 gpurun_out/${tag}_bench_pdl0.json;;
    pdl1) P2I_PDL=1 timeout 600 python bench.py --steps 100 --no-cpu --no-extras > gpurun_out/${tag}_bench_pdl1.json 2> gpurun_out/${tag}_bench_pdl1.err; echo "pdl1 rc=$?"; head -c 300 gpurun_out/${tag}_bench_pdl1.json;;
    testsel) timeout 900 python -m pytest tests -m gpu -q -rf -p no:cacheprovider -k "$P2I_TESTSEL" > gpurun_out/${tag}_pytest_sel.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_sel.txt; tail -30 gpurun_out/${tag}_pytest_sel.txt;;
    peer2) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/run_peer_allreduce.py > gpurun_out/${tag}_peer2.txt 2>&1; echo "peer2 rc=$?"; tail -15 gpurun_out/${tag}_peer2.txt;;
    bench2) for b in 1 0; do P2I_BUCKETED=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$b bench.py --gpus 2 --steps 100 --no-cpu --no-extras > gpurun_out/${tag}_bench_2gpu_bucketed$b.json 2> gpurun_out/${tag}_bench_2gpu_bucketed$b.err; echo "bench2 bucketed=$b rc=$?"; head -c 300 gpurun_out/${tag}_bench_2gpu_bucketed$b.json; echo; done;;
    graph2) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/prof_graph.py gpurun_out/${tag}_graph2_timeline.txt > gpurun_out/${tag}_graph2_kernels.txt 2>&1; echo "graph2 rc=$?"; grep "^kernels" gpurun_out/${tag}_graph2_kernels.txt;;
    blocks2) for nb in 8 16 64; do P2I_PEER_BLOCKS=$nb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957${nb:0:1} bench.py --gpus 2 --steps 100 --no-cpu --no-extras > gpurun_out/${tag}_bench_2gpu_blocks$nb.json 2> gpurun_out/${tag}_bench_2gpu_blocks$nb.err; echo "blocks=$nb rc=$?"; head -c 250 gpurun_out/${tag}_bench_2gpu_blocks$nb.json; echo; done;;
    bench8) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 400 --warmup 5 > gpurun_out/${tag}_bench_train_8gpu.json 2> gpurun_out/${tag}_bench_train_8gpu.err; echo "bench8 train rc=$?"; head -c 300 gpurun_out/${tag}_bench_train_8gpu.json; echo
            timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --workload infer --steps 400 --warmup 5 > gpurun_out/${tag}_bench_infer_8gpu.json 2> gpurun_out/${tag}_bench_infer_8gpu.err; echo "bench8 infer rc=$?"; head -c 300 gpurun_out/${tag}_bench_infer_8gpu.json; echo
            P2I_BUCKETED=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 100 --warmup 5 --no-extras --no-cpu > gpurun_out/${tag}_bench_train_8gpu_bucketed.json 2> gpurun_out/${tag}_bench_train_8gpu_bucketed.err; echo "bench8 bucketed rc=$?"; head -c 300 gpurun_out/${tag}_bench_train_8gpu_bucketed.json; echo;;
    bench4) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 400 --warmup 5 > gpurun_out/${tag}_bench_train_4gpu.json 2> gpurun_out/${tag}_bench_train_4gpu.err; echo "bench4 train rc=$?"; head -c 300 gpurun_out/${tag}_bench_train_4gpu.json; echo;;
    benchq) timeout 600 python bench.py --steps 50 --no-cpu --no-extras > gpurun_out/${tag}_bench_quick.json 2> gpurun_out/${tag}_bench_quick.err; echo "benchq rc=$?"; head -c 400 gpurun_out/${tag}_bench_quick.json;;
    others) for wl in gauge1pct stress256; do timeout 600 python bench.py --workload $wl --steps 100 --no-extras --no-cpu > gpurun_out/${tag}_bench_$wl.json 2> gpurun_out/${tag}_bench_$wl.err; echo "$wl rc=$?"; done;;
    infer) timeout 600 python bench.py --workload infer --steps 100 --no-extras > gpurun_out/${tag}_bench_infer.json 2> gpurun_out/${tag}_bench_infer.err; echo "infer rc=$?";;
    ref) timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?";;
    umma) timeout 120 tools/build/umma_rate > gpurun_out/${tag}_umma_rate.txt 2>&1; echo "umma rc=$?"; cat gpurun_out/${tag}_umma_rate.txt;;
    eager) timeout 600 python tools/gpu_eager_reference.py > gpurun_out/${tag}_gpu_eager.json 2> gpurun_out/${tag}_gpu_eager.err; echo "eager rc=$?"; cat gpurun_out/${tag}_gpu_eager.json;;
    graph) timeout 300 python tools/prof_graph.py gpurun_out/${tag}_graph_timeline.txt > gpurun_out/${tag}_graph_kernels.txt 2>&1; echo "graph rc=$?"; head -3 gpurun_out/${tag}_graph_kernels.txt;;
    convab) timeout 600 python tools/bench_conv.py > gpurun_out/${tag}_conv_ab.txt 2>&1; echo "convab rc=$?";;
    wgradab) timeout 600 python tools/bench_wgrad.py > gpurun_out/${tag}_wgrad_ab.txt 2>&1; echo "wgradab rc=$?";;
    ncunew) # full capture of the kernels changed last (d3d.0 MMA kernels, UPPos backward) inside one training step; summary only travels back
      timeout 200 ncu --set full --clock-control none --import-source on -k regex:"d3d_first|upmod_bwd" -s 24 -c 12 -o gpurun_out/${tag}_new_full \
          python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_ncu_new.log 2>&1
      echo "ncu new rc=$?"
      ncu -i gpurun_out/${tag}_new_full.ncu-rep --page raw --csv > /tmp/new_raw.csv 2>/dev/null && python tools/ncu_summarize.py --glue /tmp/new_raw.csv gpurun_out/${tag}_new_full_summary.txt
      mv gpurun_out/${tag}_new_full.ncu-rep /tmp/ 2>/dev/null; tail -14 gpurun_out/${tag}_new_full_summary.txt | cut -c1-200;;
    ncul)
      timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 420 --csv --log-file gpurun_out/${tag}_launches.csv \
          python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_ncu_list.log 2>&1
      echo "ncu list rc=$?";;
    ncu)
      timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err &&
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 420 --csv --log-file gpurun_out/${tag}_launches.csv \
          python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_ncu_list.log 2>&1
      echo "ncu list rc=$?"
      # the 35 tensor-core launches of one generator forward (all four levels: <128,false,2> <256,false,2> <128,true,2> <64,true,1>)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 330 -c 36 -o gpurun_out/${tag}_halo_full \
          python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_ncu_halo.log 2>&1
      echo "ncu halo rc=$?"
      ncu -i gpurun_out/${tag}_halo_full.ncu-rep --page raw --csv > /tmp/halo_raw.csv 2>/dev/null && python tools/ncu_summarize.py /tmp/halo_raw.csv gpurun_out/${tag}_halo_full_summary.txt
      mv gpurun_out/${tag}_halo_full.ncu-rep /tmp/ 2>/dev/null
      # weight gradients: the last discriminator ones, then the generator's level-0 / level-1 launches
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 155 -c 22 -o gpurun_out/${tag}_wgrad_full \
          python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_ncu_wgrad.log 2>&1
      echo "ncu wgrad rc=$?"
      ncu -i gpurun_out/${tag}_wgrad_full.ncu-rep --page raw --csv > /tmp/wgrad_raw.csv 2>/dev/null && python tools/ncu_summarize.py /tmp/wgrad_raw.csv gpurun_out/${tag}_wgrad_full_summary.txt
      mv gpurun_out/${tag}_wgrad_full.ncu-rep /tmp/ 2>/dev/null
      du -sh gpurun_out;;
  esac
done
