set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  $TR --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 2>gpurun_out/err_strong_n$n.log | tail -1 > gpurun_out/bench_r2h_n${n}_strong.json
done
$TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 3 --scaling weak 2>gpurun_out/err_weak_n8.log | tail -1 > gpurun_out/bench_r2h_n8_weak.json
rm -f gpurun_out/config5_sweep.jsonl
python tools/sweep_config5.py --gpus 8 --sizes 1024,8192 --out gpurun_out/config5_sweep_r2h_n8.jsonl
python tools/phase_profile.py peptide_cl 148 > gpurun_out/phase_r2h_peptide_cl.txt 2>&1
for f in gpurun_out/bench_r2h_n*_*.json; do echo $f; cut -c1-330 $f; done
