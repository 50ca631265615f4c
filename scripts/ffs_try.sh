set -u
cd gpurun_out
for tend in 0.05 0.5 3.0; do
  sed "s/set t_end = 3.000/set t_end = $tend/; s/set n_writeout_frames = 100/set n_writeout_frames = 2/" ../examples/five-moment/forward_facing_step.inp > ffs_$tend.inp
  ( time timeout 250 ../warpii_b200/bin/forward_facing_step ffs_$tend.inp ) 2>&1 | tail -6
done
ls FiveMoment__ffs_0.5 | head
