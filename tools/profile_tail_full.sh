# ncu --set full capture of the wrapper tail's image-shift kernel (Atari-like step)
sed -n "2,17p" tools/profile_tail.sh > /tmp/tail_run.py
ncu --set full --clock-control none --import-source on -k regex:tail_image_shift -s 2 -c 1 \
    -f -o gpurun_out/r2_wrapper_tail_shift python /tmp/tail_run.py > gpurun_out/r2_wrapper_tail_shift.log 2>&1
