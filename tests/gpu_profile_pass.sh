#!/bin/bash
# One GPU call: launch list, whole-frame DRAM traffic (caches NOT flushed between kernels) and --set full captures of the
# kernels that carry the frame.  Usage (on the GPU box): bash tests/gpu_profile_pass.sh TAG [full]
# Outputs under gpurun_out/: TAG_launches.csv, TAG_traffic.csv, TAG_top.ncu-rep, TAG_small.ncu-rep
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export DSPMAP_NCU=1
export DSPMAP_NORM_POLL=0   # ncu serialises kernels: the normaliser cannot wait for a kernel that runs after it
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --only-headline"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_launches.log 2>&1
timeout 600 ncu --profile-from-start off --cache-control none --clock-control none -c 400 --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
    --log-file $OUT/${TAG}_traffic.csv $B > $OUT/${TAG}_traffic.log 2>&1
if [ "$2" = "full" ]; then
  timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none \
      -k regex:'k_weight2|k_cz_chain|k_pair_eval|k_resample|k_nb_place|k_predict' -c 12 -f -o $OUT/${TAG}_top $B > $OUT/${TAG}_top.log 2>&1
  timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none \
      -k regex:'k_pyr_sort|k_nb_point1|k_arrive|k_occ_count|k_norm|k_nb_cand|k_group_scatter|k_nb_mask|k_obs_rank|k_pair_prep|k_enumerate' -c 22 -f -o $OUT/${TAG}_small $B > $OUT/${TAG}_small.log 2>&1
fi
ls -la $OUT | grep $TAG
