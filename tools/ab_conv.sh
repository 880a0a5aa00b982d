# A-B of ie_conv_set_mode flags on single layers, interleaved: tools/ab_conv.sh "<layer filter>" flagsA flagsB [reps]
for i in 1 2 3; do
  for f in $2 $3; do
    echo -n "flags=$f  "; python tools/conv_bench.py --only "$1" --flags $f --reps ${4:-5} 2>&1 | grep -v total | tr '\n' ' '; echo
  done
done
