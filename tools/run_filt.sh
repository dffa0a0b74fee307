timeout 600 python -m pytest tests -m gpu -x -q -k "pfb or channelizer or Polyphase" 2>&1 | tail -5
timeout 200 python tools/pfb_ab.py 2>&1 | head -8
