for v in main slots2; do
  LIB=$PWD/ott-packager_b200/variants/libottgpu_$v.so; [ "$v" = "main" ] && LIB=$PWD/ott-packager_b200/libottgpu.so
  for caps in "48 27" "52 29" "56 30" "48 32"; do set -- $caps
    echo "== $v src=$1 mid=$2"; OTTGPU_LIB=$LIB OTTGPU_SRC_CAP_KB=$1 OTTGPU_MID_CAP_KB=$2 timeout 200 python scripts/rung_cost.py 1280x720 1920x1080,1280x720,960x540,640x360,416x234 2>&1 | tail -2
  done
done
