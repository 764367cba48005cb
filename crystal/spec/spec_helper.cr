require "spec"
require "../../../src/ph-core" # ph-core checked out so that this shard sits in <ph-core>/ext/device/crystal
require "../src/device"
