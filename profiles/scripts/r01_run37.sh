for cfg in "6 3" "8 4" "6 3" "8 4"; do set -- $cfg
python bench.py --groups $1 --threads $2 > gpurun_out/b37.json 2> gpurun_out/b37.err
python - <<PY
import json
d=json.load(open('gpurun_out/b37.json'))
print("$1x$2", 'value',round(d['value']),'e2e',round(d['e2e']['value']))
PY
done
