cd $GRAFT_REPO_ROOT
for s in c3 c4 c5; do python scripts/peer_crossover_probe.py --shape $s 2>&1 | tail -1; done > gpurun_out/r02_peer_probe.txt
ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_aperture_peer.sum -k regex:crossover_peer -c 3 --csv --log-file gpurun_out/r02_ncu_crossover_peer_c5.csv python scripts/peer_crossover_probe.py --shape c5 --reps 1 > gpurun_out/r02_ncu_peer.log 2>&1
cat gpurun_out/r02_peer_probe.txt; grep -v "^==" gpurun_out/r02_ncu_crossover_peer_c5.csv | tail -12
