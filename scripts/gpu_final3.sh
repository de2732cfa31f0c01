cd $GRAFT_REPO_ROOT
O=gpurun_out/final3; mkdir -p $O
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 120 python examples/example_3state.py > $O/example_3state.log 2>&1; tail -4 $O/example_3state.log
timeout 120 python examples/example_random_graph.py > $O/example_random.log 2>&1; tail -3 $O/example_random.log
timeout 120 python examples/example_committor_probabilities.py > $O/example_committor.log 2>&1; tail -3 $O/example_committor.log
