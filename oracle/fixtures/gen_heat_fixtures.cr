# gen_heat_fixtures.cr -- golden vectors for the slice-arithmetic stencil (examples/heat_equation.cr).
# Separate from gen_fixtures.cr so that a compile problem in the example's arithmetic (NUM_POINTS is written
# as `LENGTH // SPACING + 1`, a Float64, in the reference) cannot block the other fixtures.  Usage:
#
#     cp oracle/fixtures/gen_heat_fixtures.cr $PH_CORE/examples/
#     cd $PH_CORE && crystal run examples/gen_heat_fixtures.cr -- /path/to/repo/tests/golden/ref_heat_fixtures.json
require "json"
require "../src/ph-core.cr"

include Phase

# ---- the 1-D example, verbatim except NUM_POINTS = 21 (examples/heat_equation.cr:5-51)
T_LEFT    =   0f64
T_RIGHT   = 100f64
T_INITIAL =  20f64
SPACING    = 0.05
TIMESTEP   = 0.01
NUM_POINTS = 21
CONDUCTIVITY  =  237
DENSITY       = 2700
SPECIFIC_HEAT =  900
COEFF = (CONDUCTIVITY * TIMESTEP) / (DENSITY * SPECIFIC_HEAT * (SPACING ** 2))

def update_temp(state) : NArray(Float64)
  temp_diff = NArray.fill(state.shape, 0f64)
  temp_diff[0] = (state[1] - state[0]) * COEFF
  temp_diff[-1] = (state[-2] - state[-1]) * COEFF
  (state[1...-1]).each_with_index do |center_temp, idx|
    temp_diff[idx + 1] = (state[idx] - 2 * center_temp + state[idx + 2]) * COEFF
  end
  return state.map_with_coord { |el, idx| el + temp_diff.get(idx) }
end

def simulate(state, duration)
  steps = (duration / TIMESTEP).to_i32 + 1
  steps.times { state = update_temp(state) }
  state
end

def bits(x : Float64) : String
  x.unsafe_as(UInt64).to_s
end

def bits(x : Float32) : String
  x.unsafe_as(UInt32).to_s
end

state = NArray.fill([NUM_POINTS], T_INITIAL)
state[0] = T_LEFT
state[-1] = T_RIGHT
after_1 = update_temp(state)
after_100 = simulate(state, 0.99)   # 100 steps
final = simulate(state, 100)        # 10 001 steps, what the example prints

# ---- the N-D rule of the device stencil written with reference operators (SURVEY.md 8(a) a-9), rank 2 and 3:
#      c = s[1...-1, ..]; d_k = (s[lo_k] - 2*c) + s[hi_k]; lap = (d_0 + d_1) + d_2; nxt[1...-1, ..] = c + lap * C
def step2d(s : NArray(T), coeff : T) : NArray(T) forall T
  c = s[1...-1, 1...-1]
  d0 = (s[0...-2, 1...-1] - c * T.new(2)) + s[2.., 1...-1]
  d1 = (s[1...-1, 0...-2] - c * T.new(2)) + s[1...-1, 2..]
  nxt = s.clone
  nxt[1...-1, 1...-1] = c + (d0 + d1) * coeff
  nxt
end

def step3d(s : NArray(T), coeff : T) : NArray(T) forall T
  c = s[1...-1, 1...-1, 1...-1]
  d0 = (s[0...-2, 1...-1, 1...-1] - c * T.new(2)) + s[2.., 1...-1, 1...-1]
  d1 = (s[1...-1, 0...-2, 1...-1] - c * T.new(2)) + s[1...-1, 2.., 1...-1]
  d2 = (s[1...-1, 1...-1, 0...-2] - c * T.new(2)) + s[1...-1, 1...-1, 2..]
  nxt = s.clone
  nxt[1...-1, 1...-1, 1...-1] = c + ((d0 + d1) + d2) * coeff
  nxt
end

g2 = NArray.build(6, 7) { |coord| ((coord[0] * 7 + coord[1]) * 37 % 101).to_f32 * 0.731_f32 }
g3 = NArray.build(5, 6, 8) { |coord| (((coord[0] * 6 + coord[1]) * 8 + coord[2]) * 53 % 97).to_f32 * 1.37_f32 }
g2_1 = step2d(g2, 0.1_f32)
g2_2 = step2d(g2_1, 0.1_f32)
g3_1 = step3d(g3, 0.1_f32)
g3_2 = step3d(g3_1, 0.1_f32)

path = ARGV[0]? || "ref_heat_fixtures.json"
File.open(path, "w") do |io|
  JSON.build(io) do |json|
    json.object do
      json.field "crystal_version", Crystal::VERSION
      json.field "coeff", bits(COEFF)
      json.field "heat1d_initial", state.to_a.map { |v| bits(v) }
      json.field "heat1d_after_1", after_1.to_a.map { |v| bits(v) }
      json.field "heat1d_after_100", after_100.to_a.map { |v| bits(v) }
      json.field "heat1d_final", final.to_a.map { |v| bits(v) }
      json.field "heat2d", {"shape" => [6, 7], "in" => g2.to_a.map { |v| bits(v) }, "step1" => g2_1.to_a.map { |v| bits(v) },
                            "step2" => g2_2.to_a.map { |v| bits(v) }}
      json.field "heat3d", {"shape" => [5, 6, 8], "in" => g3.to_a.map { |v| bits(v) }, "step1" => g3_1.to_a.map { |v| bits(v) },
                            "step2" => g3_2.to_a.map { |v| bits(v) }}
    end
  end
end
puts "wrote #{path}"
