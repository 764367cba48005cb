require "./spec_helper"

# The reference's specs for the hot path (spec/n_array_spec.cr:211-333, 446-466;
# spec/multi_writable_spec.cr:14-93; README.md:22-64), replayed on the device array.
# The C++ twin of this file, tests/cpp/device_narray_spec.cpp, is what the build image can
# compile and run; keep the two in step.
include Phase

private def stock_narr
  NArray[[0, 1, 2], [3, 4, 5]].to_device
end

describe DeviceNArray do
  describe "#unsafe_fetch_chunk" do
    it "returns the correct data for a simple chunk" do
      region = IndexRegion.new([1, 0..2..2], bound_shape: [2, 3])
      stock_narr.unsafe_fetch_chunk(region).to_host.should eq NArray[3, 5]
    end

    it "returns the correct data for a relative chunk" do
      region = IndexRegion.new([-2, -1..0], bound_shape: [2, 3])
      stock_narr.unsafe_fetch_chunk(region).to_host.should eq NArray[2, 1, 0]
    end

    it "returns the empty array for a zero-size chunk" do
      region = IndexRegion.new([0...0, 0...0], bound_shape: [2, 3])
      stock_narr.unsafe_fetch_chunk(region).shape.should eq [0, 0]
    end
  end

  describe "#unsafe_set_chunk" do
    it "correctly sets data for a simple chunk (device source)" do
      narr = stock_narr
      narr.unsafe_set_chunk(IndexRegion.new([1, 0..2..2], bound_shape: [2, 3]), NArray[6, 7].to_device)
      narr.to_host.should eq NArray[[0, 1, 2], [6, 4, 7]]
    end

    it "correctly sets data for a relative chunk (device source)" do
      narr = stock_narr
      narr.unsafe_set_chunk(IndexRegion.new([-2, -1..0], bound_shape: [2, 3]), NArray[6, 7, 8].to_device)
      narr.to_host.should eq NArray[[8, 7, 6], [3, 4, 5]]
    end

    it "correctly sets data for a simple chunk (scalar source)" do
      narr = stock_narr
      narr.unsafe_set_chunk(IndexRegion.new([1, 0..2..2], bound_shape: [2, 3]), 6)
      narr.to_host.should eq NArray[[0, 1, 2], [6, 4, 6]]
    end

    it "does not modify the array when given a zero-size chunk" do
      narr = stock_narr
      narr.unsafe_set_chunk(IndexRegion.new([0...0, 0...0], bound_shape: [2, 3]), 6)
      narr.to_host.should eq NArray[[0, 1, 2], [3, 4, 5]]
    end
  end

  describe "[]=(mask, value)" do
    mask = NArray[[true, false, true], [false, true, false]].to_device

    it "sets the correct elements (scalar source)" do
      narr = stock_narr
      narr[mask] = 6
      narr.to_host.should eq NArray[[6, 1, 6], [3, 6, 5]]
    end

    it "sets the correct elements (array source)" do
      narr = stock_narr
      narr[mask] = stock_narr + 10
      narr.to_host.should eq NArray[[10, 1, 12], [3, 14, 5]]
    end

    it "raises a DimensionError for a mask of the wrong shape" do
      expect_raises(DimensionError) { stock_narr[NArray.fill([3, 2], true).to_device] = 6 }
    end
  end

  describe "elementwise operators" do
    it "matches the README" do
      narr = NArray[[1, 0, 0], [0, 1, 0]].to_device
      narr2 = NArray[[0, 1, 2], [10, 11, 12]].to_device
      (narr + narr2).to_host.should eq NArray[[1, 1, 2], [10, 12, 12]]
      (narr * narr2).to_host.should eq NArray[[0, 0, 0], [0, 11, 0]]
      narr.get(0, 0).should eq 1
      narr[.., 1].to_host.should eq NArray[0, 1]
      narr.view(.., 1).to_narr.to_host.should eq NArray[0, 1]
      narr2.argmax.should eq({12, [1, 2]})
      narr2.slices(axis: 1).map(&.to_host).should eq [NArray[0, 10], NArray[1, 11], NArray[2, 12]]
    end

    it "keeps Crystal's number semantics" do
      (stock_narr ** 2).to_host.should eq NArray[[0, 1, 4], [9, 16, 25]]
      (10 - stock_narr).to_host.should eq NArray[[10, 9, 8], [7, 6, 5]]
      (stock_narr / 2).to_host.should eq NArray[[0.0, 0.5, 1.0], [1.5, 2.0, 2.5]]
      (stock_narr // -2).to_host.should eq NArray[[0, -1, -1], [-2, -2, -3]]
      (stock_narr % -4).to_host.should eq NArray[[0, -3, -2], [-1, 0, -3]]
      (stock_narr > 2).to_host.should eq NArray[[false, false, false], [true, true, true]]
    end

    it "raises the reference's errors" do
      expect_raises(ShapeError) { stock_narr + NArray.fill([3, 2], 1).to_device }
      expect_raises(DimensionError) { stock_narr.eq(NArray.fill([3, 2], 1).to_device) }
      expect_raises(OverflowError) { (DeviceNArray(Int32).fill([4], Int32::MAX) + 1); Device.raise_pending }
      (DeviceNArray(Int32).fill([4], Int32::MAX) &+ 1).get(0).should eq Int32::MIN
      expect_raises(DivisionByZeroError) { (stock_narr // 0); Device.raise_pending }
      expect_raises(Enumerable::EmptyError) { DeviceNArray(Float32).fill([3, 0, 2], 0f32).max }
    end

    it "raises instead of running blocks on the CPU" do
      expect_raises(DeviceBlockError) { stock_narr.map { |x| x ** 2 } }
      expect_raises(DeviceBlockError) { stock_narr.each_with(stock_narr) { |a, b| a + b } }
      expect_raises(DeviceBlockError) { stock_narr.view.process { |x| x } }
    end
  end

  describe "views" do
    it "folds a transform chain into one descriptor" do
      narr = NArray.build(2, 3, 4) { |_, i| i }.to_device
      chain = narr.view(.., ..-1.., 0..2..).permute.reverse
      chain.shape.should eq [2, 3, 2]
      copy = chain.to_narr
      copy.get(0, 0, 0).should eq narr.get(1, 0, 2)
      copy.get(1, 2, 1).should eq narr.get(0, 2, 0)
    end

    it "writes through a mutable view" do
      narr = NArray.fill([2, 3], 0).to_device
      narr.mutable_view.permute[.., ..] = NArray[[1, 2], [3, 4], [5, 6]].to_device
      narr.to_host.should eq NArray[[1, 3, 5], [2, 4, 6]]
    end
  end

  describe Heat do
    it "replays examples/heat_equation.cr" do
      coeff = (237 * 0.01) / (2700 * 900 * (0.05 ** 2))
      state = DeviceNArray(Float64).fill([21], 20.0)
      state[0] = 0.0
      state[-1] = 100.0
      final = Heat.simulate(state, coeff, 10_001, LibPhGpu::HeatMode::Example1D).to_host
      final.sum.should be_close(480.0, 1e-9)
      final.get(0).should be_close(14.381532, 1e-6)
      final.get(20).should be_close(42.473872, 1e-6)
    end
  end
end

# Twin of the "RowPipeline.map_rows" block of tests/cpp/device_narray_spec.cpp.
describe RowPipeline do
  it "partitions the rows in order, tapering the last chunk" do
    RowPipeline.row_chunks(8192_i64, 4_i64, 7).map { |(a, b)| b - a }.should eq [2048, 2048, 2048, 1024, 512, 256, 128, 64, 32, 16, 16]
    RowPipeline.row_chunks(10_i64, 3_i64, 2).should eq [{0_i64, 4_i64}, {4_i64, 8_i64}, {8_i64, 9_i64}, {9_i64, 10_i64}]
    RowPipeline.row_chunks(0_i64, 4_i64, 1).empty?.should be_true
  end

  it "computes a * b + c over pinned host operands like the resident arrays do" do
    rows, cols = 1000, 768
    a = NArray.build(rows, cols) { |_, i| ((i * 2654435761) % 2001).to_f32 / 1000 - 1 }
    c = NArray.build(rows, cols) { |_, i| ((i * 40503) % 1999).to_f32 / 999 - 1 }
    b = NArray.build(1, cols) { |_, i| ((i * 7919) % 2003).to_f32 / 1001 - 1 }
    want = (a.to_device.broadcast(LibPhGpu::Op::Mul, b.to_device) + c.to_device).to_host
    pa, pc, pb = PinnedArray(Float32).from(a), PinnedArray(Float32).from(c), PinnedArray(Float32).from(b)
    result = PinnedArray(Float32).new([rows, cols])
    pipe = RowPipeline.new(4_i64, 7)
    pipe.map_rows([pa, pc], result, [pb]) { |ins, shared| ins[0].broadcast(LibPhGpu::Op::Mul, shared[0]) + ins[1] }
    result.to_narr.should eq want
    pipe.close
  end

  it "raises a pipelined step's data-dependent errors at its synchronising end" do
    ia = PinnedArray(Int32).from(NArray.fill([64, 8], Int32::MAX))
    io = PinnedArray(Int32).new([64, 8])
    expect_raises(OverflowError) do
      RowPipeline.new(4_i64, 2).map_rows([ia], io) { |ins, _| ins[0] + 1 }
    end
  end
end

# Twin of the "NArray.concatenate / push / << / wrap" block of tests/cpp/device_narray_spec.cpp.
describe "joins" do
  it "concatenates along an axis with one copy per input" do
    a = NArray[[0, 1, 2], [3, 4, 5]].to_device
    b = NArray[[10, 11, 12]].to_device
    c = NArray[[20, 21], [22, 23]].to_device
    DeviceNArray(Int32).concatenate(a, b, axis: 0).to_host.should eq NArray[[0, 1, 2], [3, 4, 5], [10, 11, 12]]
    a.concatenate(c, axis: 1).to_host.should eq NArray[[0, 1, 2, 20, 21], [3, 4, 5, 22, 23]]
    expect_raises(DimensionError) { DeviceNArray(Int32).concatenate(a, c, axis: 0) }
    expect_raises(DimensionError) { DeviceNArray(Int32).concatenate(a, c, axis: -1) } # a negative axis excludes nothing
  end

  it "pushes in place and wraps" do
    a = NArray[[0, 1, 2], [3, 4, 5]].to_device
    p = a.clone
    (p << NArray[[10, 11, 12]].to_device).should be p
    p.to_host.should eq NArray[[0, 1, 2], [3, 4, 5], [10, 11, 12]]
    DeviceNArray(Int32).wrap(a, a).shape.should eq [2, 2, 3]
    expect_raises(DimensionError) { DeviceNArray(Int32).wrap(a, p) }
  end
end
