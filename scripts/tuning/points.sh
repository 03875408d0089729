#!/bin/bash
# a list of sweep points "H C B" through scripts/profile_point.py (GPU box):  POINTS="5 2 1048576;15 2 1048576" points.sh
cd "$(dirname "$0")/../.."
IFS=';' read -ra PTS <<< "${POINTS:-5 2 1048576;15 2 1048576;50 2 1048576;15 3 1048576;15 4 1048576;15 6 1048576;50 6 262144;5 6 1048576}"
for pt in "${PTS[@]}"; do set -- $pt
  echo -n "${TAG:-stock} H=$1 C=$2 B=$3 "; python scripts/profile_point.py --H $1 --C $2 --B $3 --reps ${REPS:-4}
done
