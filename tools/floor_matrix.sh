# experiment matrix: stage-B priority x tracer footprint x frames per MOG2 launch ("full" pipeline only)
for lib in "" libgofpid_b200_t256.so; do for prio in 1 0; do for g in 3 4 8; do
  echo -n "lib=$lib prio=$prio "; GOFPID_LIB=$lib GOFPID_B_PRIO=$prio timeout 100 python tools/stage_floor.py $g 2>&1 | grep full
done; done; done
