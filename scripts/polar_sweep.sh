#!/bin/bash
# BASELINE config 5: polar sweep of an airfoil for the VLM viscous-correction database, one warm-start chain of
# angles of attack per GPU, no communication.
#   scripts/polar_sweep.sh <conf.ini> <mesh_dir/> [n_gpus] [airfoil]
# builds the CLI if needed, runs one process per GPU and prints the merged table sorted by alpha.
set -e
HERE="$(cd "$(dirname "$0")/.." && pwd)"
CONF="$1"; MESH="$2"; N="${3:-$(nvidia-smi -L | wc -l)}"; AF="${4:-naca0012q}"
LIB="$HERE/aeroflex_b200/lib"
BIN="$HERE/aeroflex_b200/build/rans_cli"
mkdir -p "$(dirname "$BIN")"
[ -x "$BIN" ] || g++ -std=c++17 -O2 -fopenmp -I"$HERE/aeroflex_b200/host" -o "$BIN" "$HERE/aeroflex_b200/host/rans_cli.cpp" -L"$LIB" -laeroflex_rans_b200 -Wl,-rpath,"$LIB"
TMP="$(mktemp -d)"
for ((i = 0; i < N; i++)); do
  "$BIN" -i "$CONF" -m "$MESH" -a "$AF" --device "$i" --shard "$i/$N" -q > "$TMP/shard_$i.log" 2>&1 &
done
wait
echo "# alpha_deg CL CD CM"
cat "$TMP"/shard_*.log | grep '^POLAR' | sort -g -k2 | cut -d' ' -f2-
rm -rf "$TMP"
