#!/bin/sh
cd /root/repo && python -c "import __graft_entry__ as g; g.build(); print(\"built\")"
