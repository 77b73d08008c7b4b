#!/bin/bash
# Round-2 probe: is a real parasail (PyPI parasail-python / libparasail) reachable from the GPU box?
# Logs every attempt; exits 0 either way.  Output: gpurun_out/parasail_probe.log
out=gpurun_out/parasail_probe.log
mkdir -p gpurun_out
{
echo "== date: $(date -u)"
echo "== python -c 'import parasail'"
python -c 'import parasail; print(parasail.version, parasail.__file__)' 2>&1 | tail -2
echo "== pip download parasail (index)"
timeout 40 python -m pip download --no-deps -d /tmp/ps parasail 2>&1 | tail -3
echo "== pip index versions parasail"
timeout 40 python -m pip index versions parasail 2>&1 | tail -3
echo "== wheelhouse"
ls /opt/wheelhouse 2>/dev/null | grep -i -E 'parasail|edlib|spoa' || echo "no parasail/edlib/spoa wheel in /opt/wheelhouse"
echo "== filesystem search"
find / -xdev \( -iname '*parasail*' -o -iname 'libspoa*' -o -iname 'spoa' -o -iname '*edlib*' \) -not -path '/proc/*' 2>/dev/null | head -20
echo "(end of search)"
echo "== network"
timeout 10 python - <<'PY' 2>&1 | tail -2
import socket
try:
    socket.create_connection(("pypi.org", 443), timeout=5); print("pypi.org reachable")
except Exception as e:
    print("pypi.org unreachable:", e)
PY
} > $out 2>&1
cat $out
exit 0
