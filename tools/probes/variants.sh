# usage: bash tools/probes/variants.sh "<suffixes>" "<command>" : runs <command> once per tuning build csrc/build/variants/libpolcue_<suffix>.so
V=supervised-depth-estimation-from-polarized-images_b200/csrc/build/variants
echo "== default"; eval "$2"
for s in $1; do echo "== $s"; POLCUE_LIB=$PWD/$V/libpolcue_$s.so eval "$2"; done
