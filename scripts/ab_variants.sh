for v in default h8c4 h2c3 h8c5; do
  if [ $v = default ]; then L=""; else L="DIP_LIB=deep-image-prior_b200/libdip_$v.so"; fi
  env $L DIP_PROF_TIME=1 timeout 120 python scripts/profile_step.py 400 2>&1 | grep config | sed "s|^|[$v] |"
  env $L timeout 120 python scripts/hbm_breakdown.py 2>/dev/null | grep -E "k_bn_act_head|k_cat_stats|k_cat_write" | sort -k5 -n -r | head -5 | sed "s|^|[$v] |"
done
