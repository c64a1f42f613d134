timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for m in 3 11; do
  echo "lockstep mask $m"
  PIK_LOCKSTEP_MASK=$m python profiles/fetch_probe.py 65536 panda 2>&1 | grep -o "device_ms [0-9.]*"
  PIK_LOCKSTEP_MASK=$m python profiles/fetch_probe.py 65536 fetch 2>&1 | grep -o "device_ms [0-9.]*"
done
