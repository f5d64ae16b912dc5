#!/bin/bash
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
