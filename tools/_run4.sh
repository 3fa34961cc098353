for cfg in "IVF_POOL_ROWS=0 IVF_POOL_S2ROUTE=0" "IVF_POOL_ROWS_WAVES=1 IVF_POOL_ROUTE_WAVES=1" "IVF_POOL_ROWS_WAVES=2 IVF_POOL_ROUTE_WAVES=2" "IVF_POOL_ROWS_WAVES=4 IVF_POOL_ROUTE_WAVES=4" "IVF_POOL_ROWS_WAVES=8 IVF_POOL_ROUTE_WAVES=8" "IVF_POOL_ROWS_WAVES=16 IVF_POOL_ROUTE_WAVES=16"; do
  echo "== $cfg"; env $cfg python tools/pool_bench.py 8 2>&1 | tail -5
done
echo "== 32 clips default"; python tools/pool_bench.py 32 2>&1 | tail -5
echo "== 32 clips old"; IVF_POOL_ROWS=0 IVF_POOL_S2ROUTE=0 python tools/pool_bench.py 32 2>&1 | tail -5
