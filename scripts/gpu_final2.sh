set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/final2; mkdir -p $O
( time timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err ) 2> $O/bench_n1.time; echo "bench rc=$?"; cat $O/bench_n1.time
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/bench_ref.time; echo "ref rc=$?"; cat $O/bench_ref.time
echo done
