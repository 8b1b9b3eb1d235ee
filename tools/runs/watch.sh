#!/bin/bash
# usage: watch.sh <seconds> <logfile> <command...> : runs the command; if it is still alive after <seconds>, attaches cuda-gdb,
# dumps the resident kernels and the host stacks into <logfile>.gdb, then kills it.
T=$1; LOG=$2; shift 2
"$@" > $LOG 2>&1 &
PID=$!
for ((i = 0; i < T; i++)); do
  sleep 1
  kill -0 $PID 2>/dev/null || { wait $PID; echo "rc=$?" >> $LOG; exit 0; }
done
echo "HUNG after $T s: attaching" >> $LOG
timeout -k 5 120 cuda-gdb -p $PID -batch -ex "info cuda kernels" -ex "thread apply all bt 14" > $LOG.gdb 2>&1
kill -9 $PID 2>/dev/null
wait $PID 2>/dev/null
echo "rc=hung" >> $LOG
