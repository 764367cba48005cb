# gen_fixtures.cr -- emits reference-anchored golden vectors for the FLOATING-POINT half of the device
# path, which the reference's own specs leave unpinned (spec/multi_indexable/multi_indexable_tester.cr:754-760
# is empty; examples/heat_equation.cr prints nothing checkable).  Pure ph-core + Crystal stdlib.
#
# No Crystal toolchain exists in the build image, so this file has never been run there; it is committed so
# that any machine WITH Crystal >= 1.0 can close the gap:
#
#     cp oracle/fixtures/gen_fixtures.cr      $PH_CORE/examples/
#     cd $PH_CORE && crystal run examples/gen_fixtures.cr -- /path/to/repo/tests/golden/ref_fixtures.json
#
# tests/test_reference_fixtures.py loads the file whenever it exists (oracle on the CPU, CUDA path under
# `-m gpu`) and reports `xfail: no Crystal toolchain` while it does not.
#
# Every float travels as the decimal string of its bit pattern (UInt32 / UInt64), inputs included, so no
# text rounding sits between the two implementations.  A raised exception is recorded by class name.
require "json"
require "../src/ph-core.cr"

include Phase

def bits(x : Float64) : String
  x.unsafe_as(UInt64).to_s
end

def bits(x : Float32) : String
  x.unsafe_as(UInt32).to_s
end

def bits(x : Int) : String
  x.to_s
end

def outcome(&block : -> String) : String
  begin
    yield
  rescue ex : OverflowError | DivisionByZeroError | ArgumentError | Enumerable::EmptyError
    ex.class.name
  end
end

F64_SPECIALS = [0.0, -0.0, 1.0, -1.0, 0.1, 3.0000002, 7.5, -2.5, 1e-20, 1e300, -1e300, Float64::MIN_POSITIVE,
                5e-324, Float64::MAX, Float64::INFINITY, -Float64::INFINITY, Float64::NAN, 1.0 + Float64::EPSILON,
                123456.789, -0.3]
F32_SPECIALS = F64_SPECIALS.map { |v| v.to_f32 } + [Float32::MAX, Float32::MIN_POSITIVE, 1e-45_f32, 16777216_f32]
I32_SPECIALS = [0, 1, -1, 2, -2, 7, -7, 50000, -50000, Int32::MAX, Int32::MIN, Int32::MAX - 1, Int32::MIN + 1]

# every ordered pair of a special-value list
def pairs_of(list : Array(T)) : Array(Tuple(T, T)) forall T
  out = [] of Tuple(T, T)
  list.each { |x| list.each { |y| out << {x, y} } }
  out
end

# One element pair through the reference's own operator macro (src/multi_indexable.cr:931-946): two
# 1-element NArrays, the named operator, the element read back.
macro ewise_table(pairs, ops)
  begin
    table = {} of String => Array(String)
    {% for op in ops %}
      table[{{op}}] = {{pairs}}.map do |pair|
        outcome { bits((NArray[pair[0]].{{op.id}}(NArray[pair[1]])).get(0)) }
      end
    {% end %}
    table
  end
end

path = ARGV[0]? || "ref_fixtures.json"
f64_pairs = pairs_of(F64_SPECIALS)
f32_pairs = pairs_of(F32_SPECIALS)
i32_pairs = pairs_of(I32_SPECIALS)

f64_table = ewise_table(f64_pairs, ["+", "-", "*", "/", "//", "%", "**"])
f32_table = ewise_table(f32_pairs, ["+", "-", "*", "/", "//", "%", "**"])
i32_table = ewise_table(i32_pairs, ["+", "-", "*", "//", "%", "&+", "&-", "&*", "&", "|", "^", "<=>"])

# Int / Int -> Float64, Int ** small non-negative Int, Float ** Int32 (llvm.powi)
i32_div = i32_pairs.map { |pair| outcome { bits((NArray[pair[0]] / NArray[pair[1]]).get(0)) } }
pow_exps = [0, 1, 2, 3, 5, 10, 31, -1, -2]
i32_pow = I32_SPECIALS.map { |b| pow_exps.map { |e| outcome { bits((NArray[b] ** e).get(0)) } } }
powi_exps = [0, 1, 2, 3, 5, 10, -1, -2, -7, 1000, -1000]
f64_powi = F64_SPECIALS.map { |b| powi_exps.map { |e| outcome { bits((NArray[b] ** e).get(0)) } } }
f32_powi = F32_SPECIALS.map { |b| powi_exps.map { |e| outcome { bits((NArray[b] ** e).get(0)) } } }

# whole-array operators on one array per dtype (README.md:22-41 style): the same results in ONE call
whole_a = NArray.build(4, 5) { |coord| (coord[0] * 5 + coord[1]).to_f64 * 0.37 - 3.0 }
whole_b = NArray.build(4, 5) { |coord| 1.25 + (coord[1] - coord[0]).to_f64 * 0.11 }
whole = {
  "a"  => whole_a.to_a.map { |v| bits(v) },
  "b"  => whole_b.to_a.map { |v| bits(v) },
  "+"  => (whole_a + whole_b).to_a.map { |v| bits(v) },
  "*"  => (whole_a * whole_b).to_a.map { |v| bits(v) },
  "/"  => (whole_a / whole_b).to_a.map { |v| bits(v) },
  "//" => (whole_a // whole_b).to_a.map { |v| bits(v) },
  "%"  => (whole_a % whole_b).to_a.map { |v| bits(v) },
  "a*b+a" => (whole_a * whole_b + whole_a).to_a.map { |v| bits(v) },
  "2-a"   => (2.0 - whole_a).to_a.map { |v| bits(v) },
  "a>b"   => (whole_a > whole_b).to_a.map { |v| v ? "1" : "0" },
}

# Enumerable#sum / #min / #max over NArray#each (src/n_array.cr:556-564)
def reduce_case(values : Array(T)) : Hash(String, String | Array(String)) forall T
  narr = values.empty? ? NArray.fill([0], T.zero) : NArray.new(values)
  {
    "in"  => values.map { |v| bits(v) },
    "sum" => outcome { bits(narr.sum) },
    "min" => outcome { bits(narr.min) },
    "max" => outcome { bits(narr.max) },
  } of String => String | Array(String)
end

seq32 = Array(Float32).new(1000) { |i| (i.to_f32 * 0.1_f32) - 31.7_f32 }
seq64 = Array(Float64).new(1000) { |i| (i.to_f64 * 0.1) - 31.7 }
reductions = {
  "f32_seq"        => reduce_case(seq32),
  "f64_seq"        => reduce_case(seq64),
  "f32_neg_zero_first" => reduce_case([-0.0_f32, 0.0_f32, -0.0_f32]),
  "f32_pos_zero_first" => reduce_case([0.0_f32, -0.0_f32, 0.0_f32]),
  "f64_nan_inside" => reduce_case([1.0, Float64::NAN, 2.0]),
  "f64_nan_first"  => reduce_case([Float64::NAN, 1.0, 2.0]),
  "f64_inf"        => reduce_case([1.0, Float64::INFINITY, -Float64::INFINITY, 3.0]),
  "f32_empty"      => reduce_case([] of Float32),
  "i32_prefix_overflow" => reduce_case([Int32::MAX, 1, -5]),
  "i32_no_overflow"     => reduce_case([Int32::MAX, -5, 1]),
  "i32_ties"            => reduce_case([3, 9, -4, 9, -4]),
}

# first-maximum argmax idiom of README.md:56-61
tie = NArray.build(3, 4) { |coord| ((coord[0] * 4 + coord[1]) % 5).to_f32 }
argmax = [0, 0]
max = tie.get(argmax)
tie.each_with_coord do |el, coord|
  if el > max
    max = el
    argmax = coord.to_a
  end
end

# joins one step above gather / scatter: NArray.concatenate / #push / NArray.wrap (src/n_array.cr:321-344, 666-750) and
# get_chunk(coord, region_shape) (src/multi_indexable.cr:369-395); the reference's specs hold nothing for them
def join_outcome(&block : -> NArray(Int32)) : Hash(String, Array(Int32)) | String
  begin
    r = yield
    {"shape" => r.shape, "elements" => r.to_a}
  rescue ex : DimensionError | ShapeError | IndexError | ArgumentError
    ex.class.name
  end
end

j_a = NArray.build(2, 3, 4) { |_, i| i }
j_b = NArray.build(2, 1, 4) { |_, i| 100 + i }
j_c = NArray.build(2, 3, 2) { |_, i| 200 + i }
j_d = NArray.build(3, 3, 4) { |_, i| 300 + i }
joins = {
  "inputs" => {"a" => j_a.to_a, "b" => j_b.to_a, "c" => j_c.to_a, "d" => j_d.to_a},
  "concatenate(a,d,axis:0)"  => join_outcome { NArray.concatenate(j_a, j_d, axis: 0) },
  "concatenate(a,b,a,axis:1)" => join_outcome { NArray.concatenate(j_a, j_b, j_a, axis: 1) },
  "concatenate(a,c,axis:2)"  => join_outcome { NArray.concatenate(j_a, j_c, axis: 2) },
  "concatenate(a,a,axis:-1)" => join_outcome { NArray.concatenate(j_a, j_a, axis: -1) },
  "concatenate(a,c,axis:-1)" => join_outcome { NArray.concatenate(j_a, j_c, axis: -1) },
  "concatenate(a,b,axis:0)"  => join_outcome { NArray.concatenate(j_a, j_b, axis: 0) },
  "a.concatenate(b,axis:1)"  => join_outcome { j_a.concatenate(j_b, axis: 1) },
  "a.clone.push(d)"          => join_outcome { j_a.clone.push(j_d) },
  "a.clone.push(b)"          => join_outcome { j_a.clone.push(j_b) },
  "wrap(a,a)"                => join_outcome { NArray.wrap(j_a, j_a) },
  "wrap(a,b)"                => join_outcome { NArray.wrap(j_a, j_b) },
  "a.get_chunk([1,0,2],[1,3,2])" => join_outcome { j_a.get_chunk([1, 0, 2], [1, 3, 2]) },
  "a.get_chunk([0,0,0],[3,1,1])" => join_outcome { j_a.get_chunk([0, 0, 0], [3, 1, 1]) },
  "a.get_chunk([0,0],[1,1])"     => join_outcome { j_a.get_chunk([0, 0], [1, 1]) },
}

# Float#to_s as NArray#to_json / #to_yaml write it (src/n_array.cr:807-869)
text_values = [0.1, 2.0, -1.5e-7, 1e22, 5e-324, 0.30000000000000004, 1e14, 1e15, 1e16, 1e-4, 1e-5, -0.0, 123456789.125, 1e300]
text32 = [0.1_f32, 3.4e38_f32, 1e-38_f32, 16777216_f32, 1e16_f32, -2.5e-7_f32]

File.open(path, "w") do |io|
  JSON.build(io) do |json|
    json.object do
      json.field "crystal_version", Crystal::VERSION
      json.field "f64_specials", F64_SPECIALS.map { |v| bits(v) }
      json.field "f32_specials", F32_SPECIALS.map { |v| bits(v) }
      json.field "i32_specials", I32_SPECIALS
      json.field "ewise_f64", f64_table
      json.field "ewise_f32", f32_table
      json.field "ewise_i32", i32_table
      json.field "i32_div", i32_div
      json.field "pow_exps", pow_exps
      json.field "i32_pow", i32_pow
      json.field "powi_exps", powi_exps
      json.field "f64_powi", f64_powi
      json.field "f32_powi", f32_powi
      json.field "whole_f64", whole
      json.field "reductions", reductions
      json.field "argmax_first", {"values" => tie.to_a.map { |v| bits(v) }, "shape" => [3, 4], "max" => bits(max), "coord" => argmax}
      json.field "joins", joins
      json.field "float_text_f64", text_values.map { |v| [bits(v), v.to_s] }
      json.field "float_text_f32", text32.map { |v| [bits(v), v.to_s] }
      json.field "to_json_f64", NArray.new(text_values[0, 6]).to_json
      json.field "to_yaml_f64", NArray.new(text_values[0, 6]).to_yaml
    end
  end
end
puts "wrote #{path}"
