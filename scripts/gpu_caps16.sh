for caps in "68 18" "60 20" "56 24" "48 27" "40 30" "36 34" "30 36"; do set -- $caps
  echo "== src=$1 mid=$2"; OTTGPU_SRC_CAP_KB=$1 OTTGPU_MID_CAP_KB=$2 timeout 200 python scripts/rung_cost.py --p010 --src 3840x2160 --channels 8 --steps 20 3840x2160,2560x1440,1920x1080,1280x720,960x540,640x360 2>&1 | tail -1
done
