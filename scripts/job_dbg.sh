timeout 900 python -m pytest tests -m gpu -q -x -k "mega_tcgen05_consumer_full" 2>&1 | grep -E "Error|error|assert|^E" | head -20
