require "./device_indexable"

# `scalar op narr` (src/patches/number.cr maps it to `narr.map { scalar op e }`, a block): for a
# device array the operand order is passed to the kernel instead.
struct Number
  {% for pair in [{"+", "add"}, {"-", "sub"}, {"*", "mul"}, {"//", "floordiv"}, {"%", "mod"}, {"**", "pow"},
                  {"&+", "wadd"}, {"&-", "wsub"}, {"&*", "wmul"}, {"&**", "wpow"}, {"&", "and"}, {"|", "or"}, {"^", "xor"}] %}
    def {{pair[0].id}}(other : Phase::DeviceIndexable(T)) forall T
      other.scalar_on_left_{{pair[1].id}}(T.new(self))
    end
  {% end %}

  def /(other : Phase::DeviceIndexable(T)) forall T
    other.scalar_on_left_div(T.new(self))
  end

  {% for pair in [{">", "<"}, {"<", ">"}, {">=", "<="}, {"<=", ">="}] %}
    def {{pair[0].id}}(other : Phase::DeviceIndexable(T)) forall T
      other {{pair[1].id}} T.new(self)
    end
  {% end %}

  def eq(other : Phase::DeviceIndexable(T)) forall T
    other.eq(T.new(self))
  end
end
