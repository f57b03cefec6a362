for spec in "20:" "20:10,10" "19:" "19:10,9" "22:" "22:7,8,7" "22:8,7,7" "26:" "26:8,10,8" "26:9,8,9" "28:" "28:9,10,9"; do
  k=${spec%%:*}; g=${spec#*:}
  cols=4; [ $k -ge 26 ] && cols=1; [ $k -le 20 ] && cols=16
  echo "log_n=$k geom=${g:-default} cols=$cols: $(ZKB_NTT_GEOM=$g python tools/profile_run.py ntt --log-n $k --cols $cols --reps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms'],3),'ms', round(d['Melems_per_s']),'Melem/s')")"
done
